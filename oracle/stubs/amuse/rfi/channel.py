class MpiChannel(object):
    @staticmethod
    def is_multithreading_supported():
        return False
