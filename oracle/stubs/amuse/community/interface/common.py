class CommonCode(object):
    pass
