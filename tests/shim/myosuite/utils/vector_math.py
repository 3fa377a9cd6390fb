import numpy as np


def calculate_cosine(vec1, vec2):
    """MyoSuite utils/vector_math.py: cosine of the angle between two vectors (0 when either is ~0)."""
    if np.linalg.norm(vec1) < 1e-4 or np.linalg.norm(vec2) < 1e-4:
        return 0
    return np.dot(vec1, vec2) / (np.linalg.norm(vec1) * np.linalg.norm(vec2))
