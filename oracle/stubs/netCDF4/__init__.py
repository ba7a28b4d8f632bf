class Dataset(object):
    def __init__(self, *a, **k):
        raise RuntimeError("netCDF4 stub")
