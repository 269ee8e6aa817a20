"""Shape/length validators raising ``ValueError`` like the reference's
(/root/reference/boltzmann_machines/utils/testing.py:17-27)."""


def assert_shape(obj, name, desired_shape):
    got = getattr(obj, name).shape
    if tuple(got) != tuple(desired_shape):
        raise ValueError('`{0}` has invalid shape {1} != {2}'.format(name, got, desired_shape))


def assert_len(obj, name, desired_len):
    got = len(getattr(obj, name))
    if got != desired_len:
        raise ValueError('`{0}` has invalid len {1} != {2}'.format(name, got, desired_len))
