"""Speaker encoder on the B200 path (reference: models/encoder)."""
