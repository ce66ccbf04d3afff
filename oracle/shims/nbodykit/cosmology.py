class _Cosmology(object):
    def __init__(self, Om0):
        self.Om0 = Om0

# astropy/nbodykit Planck15: Om0 = 0.3075
Planck15 = _Cosmology(0.3075)
