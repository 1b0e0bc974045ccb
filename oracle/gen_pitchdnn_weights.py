"""oracle/gen_pitchdnn_weights.py — TEST INFRASTRUCTURE ONLY.
The weights of the reference's neural pitch estimator (dnn/pitchdnn.c: PitchDNN, pitchdnn_data.[ch]) belong to the separate model download that
is not in the tree.  This script builds the reference's own torch model definition (dnn/torch/neural-pitch/models.py: PitchDNN — two dense
layers on the 88 instantaneous-frequency features, two 3x3 conv2d layers on the 224 normalised cross-correlations, dense 288 -> 64, GRU 64,
dense 64 -> 192) with RANDOM, seeded weights and exports them exactly as dnn/torch/neural-pitch/export_neuralpitch_weights.py does (int8 dense
layers and GRU, float conv2d) into oracle/_ref/gen/pitchdnn_data.[ch].  Random weights pin the arithmetic of compute_pitchdnn() and of the
feature extraction around it (dnn/lpcnet_enc.c), not the quality of the pitch estimate.
Run by oracle/Makefile; needs /root/reference, so it only runs in the build container (the built .so travels)."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REF = os.environ.get('OPUS_REF', '/root/reference')
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'gen')
sys.path.insert(0, os.path.join(REF, 'dnn', 'torch', 'neural-pitch'))
sys.path.insert(0, os.path.join(REF, 'dnn', 'torch', 'weight-exchange'))
from wexchange.torch import dump_torch_weights   # noqa: E402
from wexchange.c_export import CWriter           # noqa: E402
from models import PitchDNN                       # noqa: E402

torch.manual_seed(0x717C)
np.random.seed(13)
with contextlib.redirect_stdout(io.StringIO()):
    model = PitchDNN()
with torch.no_grad():
    for name, p in model.named_parameters():   # torch's default range leaves the layers nearly linear: widen it
        if p.dim() >= 2:
            p.mul_(2.0)
        else:
            p.uniform_(-.3, .3)
os.makedirs(OUT, exist_ok=True)
writer = CWriter(os.path.join(OUT, 'pitchdnn_data'), message='seeded random weights (oracle/gen_pitchdnn_weights.py)', model_struct_name='PitchDNN')
writer.header.write('\n#include "opus_types.h"\ntypedef struct PitchDNN PitchDNN;\n')
with contextlib.redirect_stdout(io.StringIO()):
    for name, export_name in (('if_upsample.0', 'dense_if_upsampler_1'), ('if_upsample.2', 'dense_if_upsampler_2'), ('downsample.0', 'dense_downsampler'),
                              ('upsample.0', 'dense_final_upsampler')):
        dump_torch_weights(writer, model.get_submodule(name), name=export_name, verbose=False, quantize=True, scale=None)
    for name, export_name in (('conv.1', 'conv2d_1'), ('conv.4', 'conv2d_2')):
        dump_torch_weights(writer, model.get_submodule(name), name=export_name, verbose=False)
    units = max(dump_torch_weights(writer, model.get_submodule(name), export_name, verbose=False, input_sparse=False, quantize=True, scale=None,
                                   recurrent_scale=None) for name, export_name in (('GRU', 'gru_1'),))
writer.header.write('\n\n#define PITCH_DNN_MAX_RNN_UNITS %d\n\n' % units)
writer.close()
print('wrote', os.path.join(OUT, 'pitchdnn_data.c'))
