"""oracle/gen_dred_weights.py — TEST INFRASTRUCTURE ONLY.
Like gen_fargan_weights.py, for the DRED decoder (dnn/dred_rdovae_dec.c): the reference's own torch model (dnn/torch/rdovae/rdovae)
with seeded RANDOM weights, block-sparsified the way its training does it, exported with the reference's own exporter using the
layer list and flags of dnn/torch/rdovae/export_rdovae_weights.py into oracle/_ref/gen/dred_rdovae_dec_data.[ch] and (the encoder network,
dnn/dred_rdovae_enc.c) dred_rdovae_enc_data.[ch], plus the constants header that script writes.  Pins the arithmetic of the graphs, not their quality."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REF = os.environ.get('OPUS_REF', '/root/reference')
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'gen')
sys.path.insert(0, os.path.join(REF, 'dnn', 'torch', 'rdovae'))
sys.path.insert(0, os.path.join(REF, 'dnn', 'torch', 'weight-exchange'))
from rdovae import RDOVAE                               # noqa: E402
from wexchange.torch import dump_torch_weights          # noqa: E402
from wexchange.c_export import CWriter, print_vector    # noqa: E402

torch.manual_seed(0xD2ED)
rng = np.random.default_rng(11)
LATENT_DIM, STATE_DIM, FEATURE_DIM = 25, 50, 20          # the released model's sizes (DRED_LATENT_DIM 25, DRED_STATE_DIM 50)
with contextlib.redirect_stdout(io.StringIO()):
    model = RDOVAE(FEATURE_DIM, LATENT_DIM, 16, 256, 256, state_dim=STATE_DIM)
dec = model.get_submodule('core_decoder.module')


def block_sparsify(w, density):
    """zero whole 8 (outputs) x 4 (inputs) blocks, the granularity of the reference's sparse layers"""
    with torch.no_grad():
        o, i = w.shape
        keep = torch.from_numpy(rng.random((o // 8, i // 4)) < density)
        mask = keep.repeat_interleave(8, 0).repeat_interleave(4, 1).to(w.dtype)
        w.mul_(mask)


with torch.no_grad():
    for name, p in dec.named_parameters():
        if p.dim() >= 2:
            p.mul_(1.5)
for g in (dec.gru1, dec.gru2, dec.gru3, dec.gru4, dec.gru5):
    block_sparsify(g.weight_ih_l0, 0.5)
for c in (dec.conv1, dec.conv2, dec.conv3, dec.conv4, dec.conv5):
    block_sparsify(c.conv_dense.weight, 0.5)
block_sparsify(dec.output.weight, 0.5)
block_sparsify(dec.gru_init.weight, 0.5)

os.makedirs(OUT, exist_ok=True)
w = CWriter(os.path.join(OUT, 'dred_rdovae_dec_data'), message='seeded random model (oracle/gen_dred_weights.py)', model_struct_name='RDOVAEDec')
w.header.write('\n#include "dred_rdovae.h"\n#include "dred_rdovae_constants.h"\n\n')
with contextlib.redirect_stdout(io.StringIO()):
    for name, export_name, quantize in [('dense_1', 'dec_dense1', False), ('glu1.gate', 'dec_glu1', True), ('glu2.gate', 'dec_glu2', True),
                                        ('glu3.gate', 'dec_glu3', True), ('glu4.gate', 'dec_glu4', True), ('glu5.gate', 'dec_glu5', True),
                                        ('hidden_init', 'dec_hidden_init', False)]:
        dump_torch_weights(w, dec.get_submodule(name), name=export_name, verbose=True, quantize=quantize, scale=None)
    for name, export_name in [('output', 'dec_output'), ('conv1.conv_dense', 'dec_conv_dense1'), ('conv2.conv_dense', 'dec_conv_dense2'),
                              ('conv3.conv_dense', 'dec_conv_dense3'), ('conv4.conv_dense', 'dec_conv_dense4'), ('conv5.conv_dense', 'dec_conv_dense5'),
                              ('gru_init', 'dec_gru_init')]:
        dump_torch_weights(w, dec.get_submodule(name), name=export_name, verbose=True, sparse=True, quantize=True, scale=None)
    rnn = max(dump_torch_weights(w, dec.get_submodule(n), e, verbose=True, input_sparse=True, quantize=True, scale=None, recurrent_scale=None)
              for n, e in [('gru1', 'dec_gru1'), ('gru2', 'dec_gru2'), ('gru3', 'dec_gru3'), ('gru4', 'dec_gru4'), ('gru5', 'dec_gru5')])
    conv = max(dump_torch_weights(w, dec.get_submodule(n), e, verbose=True, quantize=True, scale=None)
               for n, e in [('conv1.conv', 'dec_conv1'), ('conv2.conv', 'dec_conv2'), ('conv3.conv', 'dec_conv3'), ('conv4.conv', 'dec_conv4'),
                            ('conv5.conv', 'dec_conv5')])
del w

# ---- the encoder half (dnn/dred_rdovae_enc.c), same treatment; layer list and flags of export_rdovae_weights.py:190-237 ----
enc = model.get_submodule('core_encoder.module')
with torch.no_grad():
    for name, p in enc.named_parameters():
        if p.dim() >= 2:
            p.mul_(1.5)
for g in (enc.gru1, enc.gru2, enc.gru3, enc.gru4, enc.gru5):
    block_sparsify(g.weight_ih_l0, 0.5)
for c in (enc.conv1, enc.conv2, enc.conv3, enc.conv4, enc.conv5):
    block_sparsify(c.conv_dense.weight, 0.5)
def pad_rows(layer, rows):   # export_rdovae_weights.py:161-181: the latent / state outputs are padded to a multiple of 8 rows (quantised layers)
    with torch.no_grad():
        w_, b_ = layer.weight, layer.bias
        pad = rows - w_.shape[0]
        layer.weight = torch.nn.Parameter(torch.cat([w_, torch.zeros(pad, w_.shape[1])], dim=0))
        layer.bias = torch.nn.Parameter(torch.cat([b_, torch.zeros(pad)], dim=0))


pad_rows(enc.z_dense, (LATENT_DIM + 7) // 8 * 8)
pad_rows(enc.state_dense_2, (STATE_DIM + 7) // 8 * 8)
we = CWriter(os.path.join(OUT, 'dred_rdovae_enc_data'), message='seeded random model (oracle/gen_dred_weights.py)', model_struct_name='RDOVAEEnc')
we.header.write('\n#include "dred_rdovae.h"\n#include "dred_rdovae_constants.h"\n\n')
with contextlib.redirect_stdout(io.StringIO()):
    for name, export_name, quantize in [('dense_1', 'enc_dense1', False), ('z_dense', 'enc_zdense', True), ('state_dense_1', 'gdense1', True),
                                        ('state_dense_2', 'gdense2', True)]:
        dump_torch_weights(we, enc.get_submodule(name), name=export_name, verbose=True, quantize=quantize, scale=None)
    for k in range(1, 6):
        dump_torch_weights(we, enc.get_submodule('conv%d.conv_dense' % k), name='enc_conv_dense%d' % k, verbose=True, sparse=True, quantize=True, scale=None)
    enc_rnn = max(dump_torch_weights(we, enc.get_submodule('gru%d' % k), 'enc_gru%d' % k, verbose=True, input_sparse=True, quantize=True, scale=None,
                                     recurrent_scale=None) for k in range(1, 6))
    enc_conv = max(dump_torch_weights(we, enc.get_submodule('conv%d.conv' % k), 'enc_conv%d' % k, verbose=True, quantize=True, scale=None) for k in range(1, 6))
del we
rnn, conv = max(rnn, enc_rnn), max(conv, enc_conv)
with open(os.path.join(OUT, 'dred_rdovae_constants.h'), 'w') as f:
    f.write('#ifndef DRED_RDOVAE_CONSTANTS_H\n#define DRED_RDOVAE_CONSTANTS_H\n')
    for k, v in [('DRED_NUM_FEATURES', FEATURE_DIM), ('DRED_LATENT_DIM', LATENT_DIM), ('DRED_STATE_DIM', STATE_DIM), ('DRED_PADDED_LATENT_DIM', (LATENT_DIM + 7) // 8 * 8),
                 ('DRED_PADDED_STATE_DIM', (STATE_DIM + 7) // 8 * 8), ('DRED_NUM_QUANTIZATION_LEVELS', 16), ('DRED_MAX_RNN_NEURONS', rnn),
                 ('DRED_MAX_CONV_INPUTS', conv), ('DRED_DEC_MAX_RNN_NEURONS', rnn)]:
        f.write('#define %s %d\n' % (k, v))
    f.write('#endif\n')

# ---- the statistical model of the latents (dred_rdovae_stats_data.[ch]): what dnn/dred_encoder.c's entropy coder reads.  Formulas of
# dump_statistical_model (export_rdovae_weights.py:71-112) on the model's own (seeded random, widened) quantisation embedding; every dimension is kept.
qemb = model.statistical_model.quant_embedding.weight.detach().clone()
levels = qemb.shape[0]
qemb = torch.reshape(qemb, (levels, 6, -1)) * 3.0 + torch.linspace(-1.0, 1.0, levels).reshape(levels, 1, 1) + 0.7 * torch.randn(levels, 6, qemb.numel() // (6 * levels))
sw = CWriter(os.path.join(OUT, 'dred_rdovae_stats_data'), message='seeded random model (oracle/gen_dred_weights.py)', enable_binary_blob=False)
sw.header.write('\n#include "opus_types.h"\n\n#include "dred_rdovae_constants.h"\n\n')
for name, wq in (('latent', qemb[:, :, :LATENT_DIM]), ('state', qemb[:, :, LATENT_DIM:LATENT_DIM + STATE_DIM])):
    quant_scales = torch.nn.functional.softplus(wq[:, 0, :]).numpy()
    dead_zone = 0.05 * torch.nn.functional.softplus(wq[:, 1, :]).numpy()
    r = torch.sigmoid(wq[:, 5, :]).numpy()
    p0 = torch.sigmoid(wq[:, 4, :]).numpy()
    p0 = 1 - r ** (0.5 + 0.5 * p0)
    quant_scales = quant_scales * (255. / 256. / (1e-15 + np.max(quant_scales, axis=0)))
    arrays = (('quant_scales', np.round(quant_scales * 2 ** 8).astype(np.uint16)), ('dead_zone', np.clip(np.round(dead_zone * 2 ** 8), 0, 255).astype(np.uint16)),
              ('r', np.clip(np.round(r * 2 ** 8), 0, 255).astype(np.uint8)), ('p0', np.clip(np.round(p0 * 2 ** 8), 0, 255).astype(np.uint16)))
    assert np.all(np.max(arrays[2][1], axis=0) > 0) and np.all(np.min(arrays[3][1], axis=0) < 255)   # the exporter would drop such a dimension
    with contextlib.redirect_stdout(io.StringIO()):
        for an, av in arrays:
            print_vector(sw.source, av, 'dred_%s_%s_q8' % (name, an), dtype='opus_uint8', static=False)
    n_dim = arrays[0][1].shape[1]
    sw.header.write(''.join('extern const opus_uint8 dred_%s_%s_q8[%d];\n' % (name, an, levels * n_dim) for an, _ in arrays) + '\n')
del sw
for ext in ('.h', '.c'):   # the writer appends an (empty) model struct and its init function that do not compile as C: the arrays are all there is
    path = os.path.join(OUT, 'dred_rdovae_stats_data' + ext)
    txt = open(path).read()
    cut = txt.index('struct Model') if ext == '.h' else txt.index('const WeightArray model_arrays')
    open(path, 'w').write(txt[:cut] + ('\n#endif\n' if ext == '.h' else ''))
print('wrote', os.path.join(OUT, 'dred_rdovae_dec_data.c'), 'max rnn', rnn, 'max conv inputs', conv)
