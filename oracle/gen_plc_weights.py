"""oracle/gen_plc_weights.py — TEST INFRASTRUCTURE ONLY.
The weights of the reference's neural packet-loss concealment (dnn/lpcnet_plc.c: PLCModel, plc_data.[ch]) are part of the separate
model download that is not in the tree.  This script builds the reference's own torch model definition (dnn/torch/plc/plc.py: dense 57 -> 128,
two 128-unit GRUs, dense 128 -> 20) with RANDOM, seeded weights and exports them the way the reference's exporter does
(dnn/torch/plc/export_plc.py: float dense layers, int8 GRUs) into oracle/_ref/gen/plc_data.[ch].  Random weights pin the arithmetic of the
predictor and the state handling of lpcnet_plc_conceal(), not the quality of the concealment.
Run by oracle/Makefile; needs /root/reference, so it only runs in the build container (the built .so travels)."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REF = os.environ.get('OPUS_REF', '/root/reference')
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'gen')
sys.path.insert(0, os.path.join(REF, 'dnn', 'torch', 'plc'))
sys.path.insert(0, os.path.join(REF, 'dnn', 'torch', 'weight-exchange'))
from wexchange.torch import dump_torch_weights   # noqa: E402
from wexchange.c_export import CWriter           # noqa: E402
import plc                                        # noqa: E402

torch.manual_seed(0x91C)
np.random.seed(11)
with contextlib.redirect_stdout(io.StringIO()):
    model = plc.PLC()
with torch.no_grad():
    for name, p in model.named_parameters():   # torch's default range leaves the GRUs nearly linear: widen it
        if p.dim() >= 2:
            p.mul_(2.0)
        else:
            p.uniform_(-.3, .3)
os.makedirs(OUT, exist_ok=True)
writer = CWriter(os.path.join(OUT, 'plc_data'), message='seeded random weights (oracle/gen_plc_weights.py)', model_struct_name='PLCModel', add_typedef=True)
writer.header.write('\n#include "opus_types.h"\n')
with contextlib.redirect_stdout(io.StringIO()):
    for name, export_name in (('dense_in', 'plc_dense_in'), ('dense_out', 'plc_dense_out')):
        dump_torch_weights(writer, model.get_submodule(name), name=export_name, verbose=False, quantize=False, scale=None)
    units = max(dump_torch_weights(writer, model.get_submodule(name), export_name, verbose=False, input_sparse=False, quantize=True, scale=None,
                                   recurrent_scale=None) for name, export_name in (('gru1', 'plc_gru1'), ('gru2', 'plc_gru2')))
writer.header.write('\n\n#define PLC_MAX_RNN_UNITS %d\n\n' % units)
writer.close()
print('wrote', os.path.join(OUT, 'plc_data.c'))
