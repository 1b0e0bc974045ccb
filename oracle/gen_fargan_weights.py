"""oracle/gen_fargan_weights.py — TEST INFRASTRUCTURE ONLY.
The FARGAN weights the reference ships are a separate download (dnn/download_model.sh) that is not in the tree, so the
reference's dnn/fargan.c cannot be compiled as is.  This script builds the reference's own torch model definition
(dnn/torch/fargan/fargan.py) with RANDOM, seeded weights and exports them with the reference's own exporter
(dnn/torch/weight-exchange) into oracle/_ref/gen/fargan_data.[ch] — the files the reference's build would generate from the
real checkpoint.  Random weights pin the arithmetic of the graph (layer order, state handling, quantisation), not audio quality.
Run by oracle/Makefile; needs /root/reference, so it only runs in the build container (the built .so travels)."""
import os
import sys
import contextlib
import io

import numpy as np
import torch
from torch import nn

REF = os.environ.get('OPUS_REF', '/root/reference')
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'gen')
sys.path.insert(0, os.path.join(REF, 'dnn', 'torch', 'fargan'))
sys.path.insert(0, os.path.join(REF, 'dnn', 'torch', 'weight-exchange'))
import wexchange.torch          # noqa: E402
import wexchange.c_export       # noqa: E402
import fargan                   # noqa: E402

torch.manual_seed(0xFA26A)
np.random.seed(7)
with contextlib.redirect_stdout(io.StringIO()):
    model = fargan.FARGAN(cond_size=256, gamma=0.9)
# init_weights leaves most layers at torch's small default range; widen a little so activations are not all ~0
with torch.no_grad():
    for name, p in model.named_parameters():
        if p.dim() >= 2:
            p.mul_(1.5)


def _remove_weight_norm(m):
    try:
        torch.nn.utils.remove_weight_norm(m)
    except ValueError:
        return


model.apply(_remove_weight_norm)
unquantized = ['cond_net.pembed', 'cond_net.fdense1', 'sig_net.cond_gain_dense', 'sig_net.gain_dense_out']   # as dump_fargan_weights.py
os.makedirs(OUT, exist_ok=True)
writer = wexchange.c_export.c_writer.CWriter(os.path.join(OUT, 'fargan_data'), model_struct_name='FARGAN', add_typedef=True)
with contextlib.redirect_stdout(io.StringIO()):
    for name, module in model.named_modules():
        quantize = name not in unquantized
        scale = None if quantize else 1 / 128
        cname = name.replace('.', '_')
        if isinstance(module, nn.Linear):
            wexchange.torch.dump_torch_dense_weights(writer, module, cname, quantize=quantize, scale=scale)
        elif isinstance(module, nn.Conv1d):
            wexchange.torch.dump_torch_conv1d_weights(writer, module, cname, quantize=quantize, scale=scale)
        elif isinstance(module, nn.GRU):
            wexchange.torch.dump_torch_gru_weights(writer, module, cname, quantize=quantize, scale=scale, recurrent_scale=scale)
        elif isinstance(module, nn.GRUCell):
            wexchange.torch.dump_torch_grucell_weights(writer, module, cname, quantize=quantize, scale=scale, recurrent_scale=scale)
        elif isinstance(module, nn.Embedding):
            wexchange.torch.dump_torch_embedding_weights(writer, module, cname, quantize=quantize, scale=scale)
writer.close()
print('wrote', os.path.join(OUT, 'fargan_data.c'))
