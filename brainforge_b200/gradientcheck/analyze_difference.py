"""gradientcheck/analyze_difference.py of the reference: verdict ladder and per-layer breakdown of the difference."""
import numpy as np

# relative-error ladder for FP32 arithmetic (the reference's float64 ladder is 1e-7 / 1e-5 / 1e-3)
PASS, SUSPICIOUS, FAIL = 1e-4, 1e-2, 1e-1


def get_results(er, verbose=1):
    code = 0 if er < PASS else 1 if er < SUSPICIOUS else 2 if er < FAIL else 3
    if verbose:
        shown = "({0:.1e})".format(er)
        print("Result of gradient check:")
        print(["Gradient check passed, error {} < %g" % PASS,
               "Suspicious gradients, %g < error {} < %g" % (PASS, SUSPICIOUS),
               "Gradient check failed, %g < error {} < %g" % (SUSPICIOUS, FAIL),
               "Fatal fail in gradient check, error {} > %g" % FAIL][code].format(shown))
    return code < 2


def fold_difference_matrices(network, dvec):
    """The flat difference vector cut back into one array per parameter tensor (weights, biases per trainable layer)."""
    pieces, at = [], 0
    for layer in network.layers[1:]:
        if not layer.trainable:
            continue
        for arr in (layer.weights, layer.biases):
            shape = [d for d in arr.shape if d > 1]
            pieces.append(dvec[at:at + arr.size].reshape(shape))
            at += arr.size
    return pieces


def analyze_difference_matrices(network, dvec):
    for number, d in enumerate(fold_difference_matrices(network, np.abs(dvec))):
        print("Sum of difference matrix no {0}: {1:.4e}".format(number, d.sum()))
        display_matrices(d)


def display_matrices(mats):
    try:
        from matplotlib import pyplot
    except ImportError:            # headless boxes: the sums above are the report
        return
    if mats.ndim > 2:
        for mat in mats:
            display_matrices(mat)
        return
    mats = np.atleast_2d(mats)
    pyplot.imshow(np.sqrt(mats.T if mats.shape[0] > mats.shape[1] else mats), cmap="hot")
    pyplot.show()
