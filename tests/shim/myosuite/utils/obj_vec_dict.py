import numpy as np


class ObsVecDict:
    """MyoSuite utils/obj_vec_dict.py subset: obs_dict -> flat vector in key order; (1, 1, ...) expansion for the vectorised
    reward functions."""

    def __init__(self):
        self.ordered_obs_keys = None

    def obsdict2obsvec(self, obs_dict, ordered_obs_keys):
        obsvec = np.zeros(0)
        for key in ordered_obs_keys:
            obsvec = np.concatenate([obsvec, np.asarray(obs_dict[key]).ravel()])
        return np.array([obs_dict["t"]]), obsvec

    @staticmethod
    def expand_dims(d):
        for key in d.keys():
            d[key] = np.asarray(d[key])[None, None, ...]

    @staticmethod
    def squeeze_dims(d):
        for key in d.keys():
            v = np.asarray(d[key])
            d[key] = np.squeeze(v, axis=(0, 1)) if v.ndim >= 2 else v
