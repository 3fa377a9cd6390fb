class PenTwirlRandomEnvV0:      # import-only (the pen model is absent and out of the hot path)
    DEFAULT_OBS_KEYS = []
    DEFAULT_RWD_KEYS_AND_WEIGHTS = {}
