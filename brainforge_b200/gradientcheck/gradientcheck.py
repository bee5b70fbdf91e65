"""gradientcheck/gradientcheck.py of the reference: `run(network, X, Y, epsilon, throw, display)`.

The analytic gradient is the network's own backward pass; the numerical one is a central-difference sweep over every
parameter that never leaves the device (raw_gradients.numerical_gradients).  The arithmetic under test is FP32 (or
TF32), not the reference's float64, which moves two numbers:

* step size: a 1e-5 step (the reference's default) is below what an FP32 cost resolves -- the rounding noise of two
  ~1e-7-relative costs divided by 2e-5 swamps the derivative.  The optimum of truncation (eps^2) against round-off
  (1e-7 / eps) sits near eps ~ 5e-3; steps smaller than MIN_EPSILON are raised to it (and the step actually taken
  between the two FP32 weights, not the nominal one, divides the cost difference).
* verdict thresholds: the reference's 1e-7 / 1e-5 / 1e-3 ladder (analyze_difference.get_results) becomes
  1e-4 / 1e-2 / 1e-1 -- measured relative errors of correct FP32 networks are 1e-4 (smooth MLP) .. 2e-3 (max-pool
  kinks inside a 4e-3 step); a wrong backward pass shows up at 1e-1 and above.
"""
import numpy as np

from .raw_gradients import analytical_gradients, numerical_gradients
from .analyze_difference import analyze_difference_matrices, get_results

MIN_EPSILON = 4e-3


def run(network, X=None, Y=None, epsilon=1e-5, throw=False, display=True):
    if X is None:
        X = np.random.normal(scale=0.1, size=network.input_shape)
    if Y is None:
        Y = np.random.normal(scale=0.1, size=network.output_shape)
    X, Y = np.asarray(X), np.asarray(Y)
    if X.shape == tuple(network.input_shape):      # the reference's default draws a single sample without a batch axis
        X, Y = X[None], Y[None]
    step = max(float(epsilon), MIN_EPSILON)
    if step != epsilon:
        print("gradientcheck: epsilon {:g} is below the resolution of FP32 costs, using {:g}".format(epsilon, step))
    analytic = analytical_gradients(network, X, Y)
    numeric = numerical_gradients(network, X, Y, step)
    diff = analytic - numeric
    norm = np.linalg.norm
    relative_error = norm(diff) / max(norm(numeric), norm(analytic))
    passed = get_results(relative_error)
    if display and not passed:
        analyze_difference_matrices(network, diff)
    if throw and not passed:
        raise RuntimeError("Gradient Check failed")
    return passed
