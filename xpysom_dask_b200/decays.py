"""Host-side sigma / learning-rate schedules (scalars, once per epoch).

Same three rules and names as the reference's decays.py:4-65, evaluated in
fp64 on the host; only their values cross the C ABI.  The exponential rule goes
through numpy (so it returns ``np.float64`` like the reference's), the other two
are plain Python arithmetic.
"""
from numpy import exp, log


def asymptotic_decay(val0, valN, curr_iter, max_iter):
    """val0 / (1 + 2 t / T); valN is ignored                (decays.py:4-20)."""
    return val0 / (1 + 2 * curr_iter / max_iter)


def exponential_decay(val0, valN, curr_iter, max_iter):
    """val0 * exp(-t * r), r = -ln(valN/val0)/T, or -ln(0.1)/T when valN == 0
    (decays.py:23-43)."""
    ratio = 0.1 if valN == 0 else valN / val0
    return val0 * exp(-curr_iter * (-log(ratio) / max_iter))


def linear_decay(val0, valN, curr_iter, max_iter):
    """straight line from val0 (t=0) to valN (t=T-1); constant if T == 1
    (decays.py:46-65)."""
    if max_iter == 1:
        return val0
    return val0 + (valN - val0) * curr_iter / (max_iter - 1)


DECAY_FUNCTIONS = {
    "exponential": exponential_decay,
    "asymptotic": asymptotic_decay,
    "linear": linear_decay,
}
