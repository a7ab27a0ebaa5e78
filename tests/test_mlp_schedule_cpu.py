"""The fused actor kernel (csrc/ttl_mlp.cuh) walks ONE list of tiles over all layers; its freedom from
deadlock rests on two properties of that list, checked here on the CPU by compiling the kernel's own
`Sched` for the host: (1) every (layer, m-tile, n-tile) appears exactly once; (2) in the global order in
which the clusters take tiles (round-robin: cluster c takes list positions c, c + n_clusters, ...), every
tile of layer l + 1 comes AFTER all tiles of layer l of the same m-tile -- so the earliest unfinished tile
never waits on a later one."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'


@pytest.fixture(scope='module')
def checker(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip('nvcc not available')
    exe = str(tmp_path_factory.mktemp('sched') / 'mlp_schedule_check')
    subprocess.check_call([NVCC, '-gencode', 'arch=compute_100a,code=sm_100a', '-std=c++17', '-O1',
                           '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(ROOT, 'tracktolearn_b200', 'csrc'),
                           os.path.join(ROOT, 'tests', 'csrc', 'mlp_schedule_check.cu'), '-o', exe])
    return exe


# (n_layers, n_m, group_m, n_clusters, n-tiles per layer): the production shape (50 000 rows: 196 m-tiles,
# groups of 33, 74 clusters, 4 n-tiles), the tail (one m-tile, 16 narrow tiles), uneven layers, fewer clusters
CASES = [(3, 196, 33, 74, (4, 4, 4, 1)), (3, 1, 1, 48, (16, 16, 16, 1)), (3, 8, 8, 74, (8, 8, 8, 1)),
         (4, 37, 5, 74, (2, 4, 1, 3)), (1, 20, 19, 74, (4, 1, 1, 1)), (2, 3, 1, 7, (5, 2, 1, 1)), (3, 0, 1, 74, (4, 4, 4, 1))]


@pytest.mark.parametrize('n_layers,n_m,group_m,n_clusters,nn', CASES)
def test_every_tile_once_and_inputs_first(checker, n_layers, n_m, group_m, n_clusters, nn):
    out = subprocess.check_output([checker] + [str(v) for v in (n_layers, n_m, group_m, n_clusters) + tuple(nn)]).decode()
    rows = [tuple(int(x) for x in line.split()) for line in out.splitlines()]
    tiles = {}
    for c, pos, l, m, n in rows:
        assert 0 <= l < n_layers and 0 <= m < n_m and 0 <= n < nn[l]
        assert (l, m, n) not in tiles, 'tile visited twice'
        tiles[(l, m, n)] = pos * n_clusters + c          # global list position
    assert len(tiles) == n_m * sum(nn[:n_layers])
    for (l, m, n), order in tiles.items():
        if l > 0:
            for j in range(nn[l - 1]):
                assert tiles[(l - 1, m, j)] < order, ((l, m, n), j)
    # a cluster takes its tiles in increasing list order
    last = {}
    for c, pos, l, m, n in rows:
        assert last.get(c, -1) < pos
        last[c] = pos
