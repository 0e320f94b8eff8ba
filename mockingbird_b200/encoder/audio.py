"""Encoder audio front-end (reference: models/encoder/audio.py:53-65).  SURVEY.md §8f row N2 - not built
yet: the reference computes it with librosa (absent here); callers that already hold the 40-channel mel
frames use inference.embed_utterance_frames / embed_frames_batch."""


def wav_to_mel_spectrogram(wav):
    raise NotImplementedError("encoder.audio.wav_to_mel_spectrogram (librosa mel front-end) is SURVEY.md §8f row N2; "
                              "pass mel frames to embed_utterance_frames / embed_frames_batch instead")
