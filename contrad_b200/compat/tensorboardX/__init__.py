"""Import stub for ``tensorboardX`` (absent from the image; reference utils.py:12 imports it at
module level, so every reference module that does ``from utils import ...`` needs the name).
Scalars are appended to ``<logdir>/scalars.tsv`` so a run still leaves a record."""
import os


class SummaryWriter(object):
    def __init__(self, logdir=None, **kwargs):
        self.logdir = logdir
        self._fh = None
        if logdir is not None:
            os.makedirs(logdir, exist_ok=True)
            self._fh = open(os.path.join(logdir, "scalars.tsv"), "a")

    def add_scalar(self, tag, value, step=None, **kwargs):
        if self._fh is not None:
            self._fh.write("%s\t%s\t%r\n" % (tag, step, float(value)))
            self._fh.flush()

    def add_image(self, *args, **kwargs):
        pass

    def add_histogram(self, *args, **kwargs):
        pass

    def flush(self):
        if self._fh is not None:
            self._fh.flush()

    def close(self):
        if self._fh is not None:
            self._fh.close()
            self._fh = None
