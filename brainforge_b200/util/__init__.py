"""Host-side helpers kept from the reference's `util` package (initialisers, batch streaming)."""
import numpy as np

floatX = "float64"   # the host API speaks float64 like the reference (config/numeric.py:1)


def white(*dims):
    """util/typing.py:22-33.  Draws from the GLOBAL NumPy RNG exactly like the reference, so the same
    seed gives the same initial weights on both back-ends."""
    if len(dims) == 2:
        fanin, fanout = dims
        return np.random.randn(fanin, fanout) * np.sqrt(2. / float(fanin + fanout))
    nf, fc, fy, fx = dims
    return np.random.randn(nf, fc, fy, fx) * np.sqrt(2. / float(nf * fy * fx + fc * fy * fx))


def batch_stream(*arrays, m, shuffle=True, infinite=True):
    """util/_nnutil.py:4-15."""
    N = arrays[0].shape[0]
    while 1:
        arg = np.arange(N)
        if shuffle:
            np.random.shuffle(arg)
        shuffled = tuple(map(lambda ary: ary[arg], arrays))
        for start in range(0, N, m):
            yield tuple(map(lambda ary: ary[start:start + m], shuffled))
        if not infinite:
            break


class MetricLogs:
    """util/logging.py:5-49 (running means per epoch; the reference's progress printing is kept minimal)."""

    def __init__(self, fields, max_steps=-1):
        self._metrics = {field: [] for field in fields}
        self.max_steps = max_steps
        self._step = 0
        self.reduced = False

    @classmethod
    def from_metric_list(cls, max_steps, cost_names=("cost",), metrics=()):
        fields = list(cost_names) + [str(metric).lower() for metric in metrics]
        return cls(fields, max_steps)

    def record(self, data):
        for key, value in data.items():
            self._metrics.setdefault(key, []).append(float(value))
        self._step += 1

    def update(self, other):
        for key, value in other._metrics.items():
            self._metrics.setdefault(key, []).extend(value)

    def mean(self, key=None):
        if key is not None:
            return float(np.mean(self._metrics[key])) if self._metrics[key] else float("nan")
        return {k: self.mean(k) for k in self._metrics}

    def reduce_mean(self):
        self._metrics = {k: [self.mean(k)] for k in self._metrics}
        self.reduced = True

    def log(self, prefix="", suffix="", add_progress=False, **print_kwargs):
        body = " ".join("{}: {:.4f}".format(k, self.mean(k)) for k in self._metrics)
        prog = " {}/{}".format(self._step, self.max_steps) if add_progress and self.max_steps > 0 else ""
        print(prefix + prog + " " + body + suffix, **print_kwargs)

    def __getitem__(self, item):
        return self._metrics[item]
