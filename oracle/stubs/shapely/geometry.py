class Point(object):
    def __init__(self, x, y=None):
        if y is None:
            x, y = x
        self.x, self.y = x, y


class Polygon(object):
    def __init__(self, *a, **k):
        pass

    def contains(self, p):
        return False


def box(*a, **k):
    return Polygon()


def shape(*a, **k):
    return Polygon()
