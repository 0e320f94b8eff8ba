"""Data hyper-parameters the inference path reads (reference: models/encoder/params_data.py:3-13)."""
mel_window_length = 25  # ms
mel_window_step = 10    # ms
mel_n_channels = 40
sampling_rate = 16000
partials_n_frames = 160
inference_n_frames = 80
