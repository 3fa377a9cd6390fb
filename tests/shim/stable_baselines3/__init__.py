"""TEST INFRASTRUCTURE: import-only stand-in (the reference's baoding.py imports VecNormalize / DummyVecEnv at module top
for its MixtureModelBaodingEnv, which the fixtures do not exercise)."""
