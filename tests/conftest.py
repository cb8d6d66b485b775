import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def yy():
    import yy_b200  # noqa: F401
    import yinyang_game_alphazero_b200 as pkg
    pkg.build()
    return pkg


@pytest.fixture(scope="session")
def host_rules_lib():
    """g++ build of the product's bitboard header for the host (test-only)."""
    import ctypes
    d = os.path.join(ROOT, "tests", "_host")
    so = os.path.join(d, "librules_host.so")
    src = os.path.join(d, "rules_host_shim.cpp")
    hdrs = [os.path.join(ROOT, "yinyang-game-alphazero_b200", "csrc", h) for h in ("yy_rules.cuh", "yy_rules_sq.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src, *hdrs]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src], check=True)
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def host_tower_lib():
    """nvcc host build of the tower's layout algebra (yy_tower.cuh) for the structural check (test-only, no GPU needed)."""
    import ctypes
    d = os.path.join(ROOT, "tests", "_host")
    so = os.path.join(d, "libtower_host.so")
    src = os.path.join(d, "tower_layout_check.cu")
    hdr = os.path.join(ROOT, "yinyang-game-alphazero_b200", "csrc", "yy_tower.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.run([nvcc, "-O1", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177",
                        "-Wno-deprecated-gpu-targets", "-o", so, src], check=True)
    return ctypes.CDLL(so)


def golden_files(prefix):
    return sorted(f for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def random_play_boards(oracle, n, m, count, seed, max_frac=0.95):
    """Legal random-play positions generated with the oracle's rules (int8[count,n,m], players)."""
    rng = np.random.default_rng(seed)
    boards = np.zeros((count, n, m), dtype=np.int8)
    players = np.ones(count, dtype=np.int8)
    plies = rng.integers(0, int(n * m * max_frac) + 1, size=count)
    for k in range(int(plies.max()) if count else 0):
        live = np.flatnonzero(plies > k)
        if live.size == 0:
            break
        masks = oracle.legal_mask(boards[live], players[live], n, m)
        acts = np.full(live.size, -1, dtype=np.int32)
        for j in range(live.size):
            idx = np.flatnonzero(masks[j])
            if idx.size:
                acts[j] = rng.choice(idx)
        nb, npl = oracle.next_state(boards[live], players[live], acts, n, m)
        boards[live], players[live] = nb, npl
    return boards, players


def randomise_bn(net, seed=1):
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.2)
                mod.running_var.copy_(torch.rand(mod.num_features, generator=g) * 1.5 + 0.25)
                mod.weight.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
            if isinstance(mod, (torch.nn.Conv2d, torch.nn.Linear)):
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.05)
    return net.eval()
