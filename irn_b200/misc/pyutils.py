"""The slice of the reference's ``misc/pyutils.py`` the label-generation steps use."""
import sys
import time

import numpy as np


class Logger(object):
    """Tee stdout to a file (misc/pyutils.py:6-17)."""

    def __init__(self, outfile):
        self.terminal = sys.stdout
        self.log = open(outfile, "w")
        sys.stdout = self

    def write(self, message):
        self.terminal.write(message)
        self.log.write(message)

    def flush(self):
        self.terminal.flush()


class Timer:
    """Wall-clock stamps per step (misc/pyutils.py:50-83, reduced to what run_sample prints)."""

    def __init__(self, starting_msg=None):
        self.start = time.time()
        if starting_msg is not None:
            print(starting_msg, time.ctime(time.time()))

    def elapsed(self):
        return time.time() - self.start


def to_one_hot(sparse_integers, maximum_val=None, dtype=bool):
    """misc/pyutils.py:86-101: [..] int -> [maximum_val, ..] one-hot."""
    a = np.asarray(sparse_integers)
    n = int(a.max()) + 1 if maximum_val is None else int(maximum_val)
    return (np.arange(n).reshape((n,) + (1,) * a.ndim) == a[None]).astype(dtype)
