"""Scratch directory of one Finder job.

The reference keeps two side files of a job in ./<uuid>/ below the current directory (nrpcalc/base/projector.py:7-19;
seq_list.txt is re-read for the output dict, repeat_graph.txt restores the graph after each vertex-cover round) and
removes the directory when the job ends.  Here the directory is a context manager: it disappears when the job leaves the
`with` block, whether it returned or raised, and an atexit hook sweeps whatever an interrupted interpreter left behind.
"""
import atexit
import os
import shutil

_live = set()


class ScratchDir(object):
    """`with ScratchDir(name) as root:` creates ./<name>/ and removes it on the way out."""

    def __init__(self, name):
        self.root = os.path.join('.', name)

    def __enter__(self):
        os.makedirs(self.root, exist_ok=True)
        _live.add(os.path.abspath(self.root))
        return self.root

    def __exit__(self, exc_type, exc, tb):
        path = os.path.abspath(self.root)
        shutil.rmtree(path, ignore_errors=True)
        _live.discard(path)
        return False

    def file(self, name):
        return '{}/{}'.format(self.root, name)


@atexit.register
def _sweep():
    for path in list(_live):
        shutil.rmtree(path, ignore_errors=True)
        _live.discard(path)
