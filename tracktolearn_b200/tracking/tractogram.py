"""Minimal ragged tractogram container (stands in for nibabel.streamlines.Tractogram, which
the reference returns from get_streamlines, tracking_env.py:289-294; nibabel is optional)."""
import numpy as np


class TractogramItem(object):
    def __init__(self, streamline, data_for_streamline, data_for_points=None):
        self.streamline = streamline
        self.data_for_streamline = data_for_streamline
        self.data_for_points = data_for_points or {}


class Tractogram(object):
    """Streamlines stored packed: ``data`` [sum(L),3] float32 + ``offsets`` [n+1] int64."""

    def __init__(self, data=None, offsets=None, data_per_streamline=None, streamlines=None,
                 affine_to_rasmm=None):
        if streamlines is not None:
            lens = np.asarray([len(s) for s in streamlines], dtype=np.int64)
            offsets = np.concatenate(([0], np.cumsum(lens)))
            data = (np.concatenate(streamlines).astype(np.float32) if len(streamlines)
                    else np.zeros((0, 3), np.float32))
        self.data = data if data is not None else np.zeros((0, 3), np.float32)
        self.offsets = offsets if offsets is not None else np.zeros((1,), np.int64)
        self.data_per_streamline = data_per_streamline or {}
        self.affine_to_rasmm = affine_to_rasmm

    @property
    def streamlines(self):
        return [self.data[self.offsets[i]:self.offsets[i + 1]] for i in range(len(self))]

    @property
    def lengths(self):
        return np.diff(self.offsets)

    def __len__(self):
        return len(self.offsets) - 1

    def __iter__(self):
        for i in range(len(self)):
            yield TractogramItem(self.data[self.offsets[i]:self.offsets[i + 1]],
                                 {k: v[i] for k, v in self.data_per_streamline.items()})

    def __iadd__(self, other):
        self.data = np.concatenate((self.data, other.data))
        self.offsets = np.concatenate((self.offsets, other.offsets[1:] + self.offsets[-1]))
        for k in self.data_per_streamline:
            self.data_per_streamline[k] = np.concatenate(
                (self.data_per_streamline[k], other.data_per_streamline[k]))
        return self
