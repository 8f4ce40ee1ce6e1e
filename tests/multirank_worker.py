"""world_size-2 worker (gloo, CPU): shards shuffle blocks i -> rank i mod N, each rank produces per-block metadata with
the ORACLE standing in for the codec (no GPU here), gathers it, and checks the assembled .index/.checksum equal the
single-process result.  Launched by tests/test_multirank.py through torch.distributed.run."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spark_s3_shuffle_b200 as pkg  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    rk = pkg.ranks.Ranks(backend="gloo")
    assert rk.world == 2
    n_blocks, recs = 11, 300  # odd count: ranks own 6 and 5 blocks
    blocks = rk.my_blocks(n_blocks)
    assert list(blocks) == list(range(rk.rank, n_blocks, 2))
    mine = []
    for b in blocks:
        raw = oracle.gen_terasort(int(b) * recs, recs).tobytes()
        comp = oracle.lz4block_compress(raw, 32768)
        mine.append((len(comp), oracle.crc32c(comp)))
    rk.barrier()
    full = rk.gather_block_results(n_blocks, np.array(mine, dtype=np.int64))
    t = rk.max_over_ranks(1.0 + rk.rank)
    s = rk.sum_over_ranks(len(blocks))
    if rk.rank == 0:
        ref = []
        for b in range(n_blocks):
            comp = oracle.lz4block_compress(oracle.gen_terasort(b * recs, recs).tobytes(), 32768)
            ref.append((len(comp), oracle.crc32c(comp)))
        ok = bool((full == np.array(ref, dtype=np.int64)).all())
        index = oracle.index_bytes(full[:, 0])
        print(json.dumps({"ok": ok, "max": t, "sum": s, "index_len": len(index),
                          "same_index": index == oracle.index_bytes([r[0] for r in ref])}))
    rk.close()


if __name__ == "__main__":
    main()
