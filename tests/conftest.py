import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_PARENT = os.path.join(ROOT, 'boltzmann-machines_b200')
for p in (ROOT, PKG_PARENT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture
def oracle_engines():
    """Route the package's models to the numpy oracle (host-logic tests on CPU)."""
    from boltzmann_machines.base import set_engine_factory
    from oracle.rbm import rbm_factory
    old = set_engine_factory('rbm', rbm_factory)
    try:
        from oracle.dbm import dbm_factory
        old_d = set_engine_factory('dbm', dbm_factory)
    except ImportError:
        old_d = None
    yield
    set_engine_factory('rbm', old)
    set_engine_factory('dbm', old_d)


@pytest.fixture(params=['oracle',
                        pytest.param('cuda-fp32', marks=pytest.mark.gpu),
                        pytest.param('cuda-bf16', marks=pytest.mark.gpu)])
def engines(request, monkeypatch):
    """The engine behind the package's models: numpy oracle (CPU) or libbm.so (GPU)."""
    from boltzmann_machines.base import set_engine_factory
    if request.param == 'oracle':
        from oracle.rbm import rbm_factory
        from oracle.dbm import dbm_factory
        old = set_engine_factory('rbm', rbm_factory)
        old_d = set_engine_factory('dbm', dbm_factory)
        yield request.param
        set_engine_factory('rbm', old)
        set_engine_factory('dbm', old_d)
    else:
        monkeypatch.setenv('BM_COMPUTE', request.param.split('-')[1])
        old = set_engine_factory('rbm', None)
        old_d = set_engine_factory('dbm', None)
        yield request.param
        set_engine_factory('rbm', old)
        set_engine_factory('dbm', old_d)


@pytest.fixture
def workdir(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    return tmp_path
