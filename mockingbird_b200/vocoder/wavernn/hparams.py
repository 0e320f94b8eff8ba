"""WaveRNN hyper-parameters (reference: models/vocoder/wavernn/hparams.py, which re-exports the
synthesizer's audio settings models/synthesizer/hparams.py:5-15).  Note the reference's own
inconsistency, reproduced on purpose: hop_length is the synthesizer's 256 while
the upsample factors multiply to 200, so generate() never trims and fades out over 5120 samples."""

# Audio settings (models/synthesizer/hparams.py)
sample_rate = 16000
n_fft = 1024
num_mels = 80
hop_length = 256
win_length = 1024
fmin = 55
min_level_db = -100
ref_level_db = 20
mel_max_abs_value = 4.0
preemphasis = 0.97
apply_preemphasis = True

bits = 9
mu_law = True

voc_mode = 'RAW'
voc_upsample_factors = (5, 5, 8)
voc_rnn_dims = 512
voc_fc_dims = 512
voc_compute_dims = 128
voc_res_out_dims = 128
voc_res_blocks = 10
voc_pad = 2

voc_gen_batched = True
voc_target = 8000
voc_overlap = 400
