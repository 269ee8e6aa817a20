"""Naming convention that separates constructor parameters from fitted
attributes (/root/reference/boltzmann_machines/base/base.py:1-5):
``name`` -> parameter, ``name_`` -> attribute, ``_name`` -> private."""


def is_param_name(name):
    return not (name.startswith('_') or name.endswith('_'))


def is_attribute_name(name):
    return name.endswith('_') and not name.startswith('_')
