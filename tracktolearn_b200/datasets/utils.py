"""Volume container + mm->voxel conversion (reference: datasets/utils.py:10-124)."""
import numpy as np


class MRIDataVolume(object):
    """Data volume with its vox->rasmm affine (reference: datasets/utils.py:10-45)."""

    def __init__(self, data=None, affine_vox2rasmm=None):
        self._data = data
        self.affine_vox2rasmm = affine_vox2rasmm

    @property
    def data(self):
        return self._data

    @property
    def shape(self):
        return self.data.shape


def convert_length_mm2vox(length_mm, affine_vox2rasmm):
    """Reference: datasets/utils.py:88-124 (same ValueError on non-isotropic voxels)."""
    diag = np.diagonal(affine_vox2rasmm)[:3]
    vox2mm = np.mean(np.abs(diag))
    if not np.allclose(np.abs(diag), vox2mm, rtol=5e-2, atol=5e-2):
        raise ValueError("Voxel space is not iso, "
                         " cannot convert a scalar length "
                         "in mm to voxel space. "
                         "Affine provided : {}".format(affine_vox2rasmm))
    return length_mm / vox2mm


def get_sh_order_and_fullness(ncoeffs):
    """scilpy.reconst.utils.get_sh_order_and_fullness restated: symmetric bases have
    (o+1)(o+2)/2 coefficients, full bases (o+1)^2."""
    for order in range(0, 34, 2):
        if (order + 1) * (order + 2) // 2 == ncoeffs:
            return order, False
    for order in range(0, 34):
        if (order + 1) ** 2 == ncoeffs:
            return order, True
    raise ValueError('Invalid number of coefficients: {}'.format(ncoeffs))
