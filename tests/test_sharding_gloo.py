"""CPU, world_size 2 (gloo): the multi-rank bookkeeping bench.py uses — contiguous batch sharding with no
data-path collective, barrier + max-reduction of the elapsed time, rank 0 reporting whole-job throughput."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_sharding_and_max_timing(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        import bench
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        lo, hi = bench.shard(2048, w, r)
        # every rank "processes" its shard; no tensor crosses ranks on the data path
        n = torch.tensor([hi - lo], dtype=torch.int64)
        t = torch.tensor([1.0 + r], dtype=torch.float64)          # rank 1 is slower
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)                  # bookkeeping only (off the hot path)
        if r == 0:
            print(json.dumps({{"images": int(n.item()), "ms": float(t.item()), "lo": lo, "hi": hi}}))
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d == {"images": 2048, "ms": 2.0, "lo": 0, "hi": 1024}
