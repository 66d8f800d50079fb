import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "stress-particle-sph_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    import __graft_entry__
    __graft_entry__.build()


@pytest.fixture(scope="session")
def deck_dir(tmp_path_factory):
    """directory factory: deck_dir('bui') -> path with the regenerated shipped deck"""
    from spsph import decks
    cache = {}

    def make(kind, **kw):
        key = (kind, tuple(sorted(kw.items())))
        if key not in cache:
            d = tmp_path_factory.mktemp(f"deck_{kind}")
            decks.write_deck(str(d), decks.SHIPPED[kind](**kw) if kind in decks.SHIPPED else getattr(decks, kind)(**kw))
            cache[key] = str(d)
        return cache[key]
    return make
