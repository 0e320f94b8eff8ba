"""Data hyper-parameters the inference path reads (reference: models/encoder/params_data.py:3-13)."""
mel_window_length = 25  # ms
mel_window_step = 10    # ms
mel_n_channels = 40
sampling_rate = 16000
partials_n_frames = 160
inference_n_frames = 80

# audio volume normalisation / VAD (models/encoder/params_data.py:16-29), read by preprocess_wav
vad_window_length = 30  # ms
vad_moving_average_width = 8
vad_max_silence_length = 6
audio_norm_target_dBFS = -30
