"""Model hyper-parameters of the speaker encoder (reference: models/encoder/params_model.py:3-5)."""
model_hidden_size = 256
model_embedding_size = 256
model_num_layers = 3
