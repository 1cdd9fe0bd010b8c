"""Host-side scalar helpers shared by the class API and the batch drivers (pure numpy; no spectrum arithmetic)."""
import numpy as np


def risco(a):
    """Bardeen ISCO (relagn.py:583-596)."""
    a = np.asarray(a, dtype=np.float64)
    Z1 = 1 + (1 - a ** 2) ** (1 / 3) * ((1 + a) ** (1 / 3) + (1 - a) ** (1 / 3))
    Z2 = np.sqrt(3 * a ** 2 + Z1 ** 2)
    return 3 + Z2 - np.sign(a) * np.sqrt((3 - Z1) * (3 + Z1 + 2 * Z2))
