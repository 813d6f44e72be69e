"""Stand-in for torchio.Subject (test infrastructure only); drr.py:18,64,73,83,86."""


class _Image:
    def __init__(self, data, affine):
        self.data = data
        self.affine = affine


class Subject:
    def __init__(self, volume, affine, reorient, mask=None):
        self.volume = _Image(volume, affine)
        self.density = _Image(volume, affine)
        self.mask = None if mask is None else _Image(mask, affine)
        self.reorient = reorient
        self.orientation = "AP"
        self.fiducials = None
