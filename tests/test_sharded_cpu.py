"""world_size-2 and -3 gloo tests of the multi-rank host logic on CPU (no GPU needed): sharded
columns (memfd + mmap standing in for cuMemCreate + cuMemMap), SCM_RIGHTS descriptor exchange,
the controller protocol and the per-tick barrier.  SURVEY §8e invariant: the final state is
identical for every number of shards, and identical to the oracle's."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3, 4])
def test_sharded_host_logic_gloo(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29540 + world),
           os.path.join(ROOT, "tests", "sharded_worker_cpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ok"] and res["world"] == world
