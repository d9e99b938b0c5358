"""Four-line stand-in so that the reference package (baseline/_ref) imports in images
without the real `itsdangerous`; only praline's RemoteManager would use it."""


class BadSignature(Exception):
    pass


class Serializer(object):
    def __init__(self, *a, **k):
        pass
