"""Iteration helpers and log-domain statistics used by the training loops and
by AIS post-processing.  Same names and behaviour as
/root/reference/boltzmann_machines/utils/utils.py:10-170 (its doctest values
are reproduced in tests/test_utils.py); the implementations are our own.
"""
import numpy as np

try:                                     # progress bars are optional garnish
    from tqdm import tqdm as _tqdm
except Exception:                        # pragma: no cover
    _tqdm = None


def write_during_training(s):
    """Print a line without breaking an active progress bar."""
    if _tqdm is not None:
        _tqdm.write(s)
    else:
        print(s)


def _maybe_bar(it, verbose, **kw):
    if verbose and _tqdm is not None:
        return _tqdm(it, **kw)
    return it


def batch_iter(X, batch_size=10, verbose=False, desc='epoch'):
    """Yield consecutive row slices of ``X``; the last one may be short.

    >>> [b.tolist() for b in batch_iter(np.arange(10).reshape(5, 2), batch_size=2)]
    [[[0, 1], [2, 3]], [[4, 5], [6, 7]], [[8, 9]]]
    """
    X = np.asarray(X)
    starts = range(0, len(X), batch_size)
    for s in _maybe_bar(starts, verbose, leave=False, ncols=64, desc=desc):
        yield X[s:s + batch_size]


def batch_bounds(n_rows, batch_size):
    """(start, stop) pairs of the slices ``batch_iter`` would produce."""
    return [(s, min(s + batch_size, n_rows)) for s in range(0, n_rows, batch_size)]


def epoch_iter(start_epoch, max_epoch, verbose=False):
    """1-based epoch numbers ``start_epoch+1 .. max_epoch``."""
    for e in _maybe_bar(range(start_epoch + 1, max_epoch + 1), verbose,
                        leave=True, ncols=84, desc='training'):
        yield e


def make_list_from(x):
    """Scalar -> one-element list; iterable -> list."""
    return list(x) if hasattr(x, '__iter__') else [x]


def one_hot(y, n_classes=None):
    """
    >>> one_hot([2, 0]).tolist()
    [[0.0, 0.0, 1.0], [1.0, 0.0, 0.0]]
    """
    y = np.asarray(y, dtype=int)
    n_classes = n_classes or int(y.max()) + 1
    out = np.zeros((len(y), n_classes))
    out[np.arange(len(y)), y] = 1.
    return out


def one_hot_decision_function(y):
    """Row-wise arg-max as a one-hot matrix.

    >>> one_hot_decision_function([[0.1, 0.4, 0.5], [0.8, 0.1, 0.1]]).tolist()
    [[0.0, 0.0, 1.0], [1.0, 0.0, 0.0]]
    """
    y = np.asarray(y)
    out = np.zeros_like(y, dtype=float)
    out[np.arange(len(y)), y.argmax(axis=1)] = 1.
    return out


def unhot(y, n_classes=None):
    """
    >>> unhot([[0, 0, 1], [0, 1, 0]]).tolist()
    [2, 1]
    """
    y = np.asarray(y)
    n_classes = n_classes or y.shape[1]
    return y.dot(np.arange(n_classes))


def log_sum_exp(x):
    """log(sum(exp(x))), shift-stabilised.

    >>> round(float(log_sum_exp([1000, 1001, 1000])), 3)
    1001.551
    """
    x = np.asarray(x, dtype=float)
    m = x.max()
    return m + np.log(np.exp(x - m).sum())


def log_mean_exp(x):
    """log(mean(exp(x))).

    >>> round(float(log_mean_exp([1, 2, 3])), 4)
    2.309
    """
    return log_sum_exp(x) - np.log(len(x))


def log_diff_exp(x):
    """log(diff(exp(x))) for an increasing sequence.

    >>> np.round(log_diff_exp([1, 2, 3]), 4).tolist()
    [1.5413, 2.5413]
    """
    x = np.asarray(x, dtype=float)
    m = x.max()
    return m + np.log(np.diff(np.exp(x - m)))


def log_std_exp(x, log_mean_exp_x=None):
    """log(std(exp(x))) via log E[e^{2x}] and log E[e^x].

    >>> round(float(log_std_exp(np.arange(8.))), 6)
    5.875416
    """
    x = np.asarray(x, dtype=float)
    m = log_mean_exp(x) if log_mean_exp_x is None else log_mean_exp_x
    return 0.5 * log_diff_exp([2. * m, log_mean_exp(2. * x)])[0]
