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
def _built(tmp_path_factory):
    import __graft_entry__
    __graft_entry__.build()
    if os.environ.get("SPSPH_EMULATE"):
        # SPSPH_EMULATE=1 python -m pytest tests -m gpu: the GPU suite on the HOST-EMULATED engine (no GPU needed; see
        # tests/test_step_emulation_cpu.py). spsph.Engine then binds the emulated build of csrc/spsph_engine.cu.
        import subprocess
        d = tmp_path_factory.mktemp("emu_engine_session")
        cpp, so = str(d / "engine_host.cpp"), str(d / "libspsph_emu.so")
        subprocess.run([sys.executable, os.path.join(ROOT, "tests", "native", "make_engine_host.py"),
                        os.path.join(ROOT, "stress-particle-sph_b200", "csrc", "spsph_engine.cu"), cpp], check=True,
                       stdout=subprocess.DEVNULL)
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared", "-w",
                        "-D__noinline__=", "-fno-gnu-unique", "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "tests", "native"),
                        "-I" + os.path.join(ROOT, "stress-particle-sph_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
                        "-o", so, cpp, "-ldl"], check=True)
        import spsph.engine as E
        E._lib, E._CUDA_SO = None, so


def pytest_collection_modifyitems(config, items):
    if not os.environ.get("SPSPH_EMULATE"):
        return
    skip = pytest.mark.skip(reason="not available on the serial host-emulated engine (multi-GPU, 4 M particles, driver "
                                   "binary; the tile kernels' warp collectives need the SIMT emulation: "
                                   "tests/test_step_emulation_cpu.py::test_emulated_engine_simt_mode_tile_path)")
    for it in items:
        if any(k in it.nodeid for k in ("test_multi_gpu", "4m_bitwise", "driver_frames", "bui_full", "test_gpu_tile_path")):
            it.add_marker(skip)


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
