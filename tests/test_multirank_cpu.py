"""CPU, world_size 2, gloo: the N>1 host logic — weight broadcast, stream sharding, max-over-ranks timing — with the
oracle standing in for the device (no GPU here).  Sharded results must equal the unsharded run bit for bit: the path
has no cross-stream dependency, which is exactly why it needs no data-path collective."""
import hashlib, os, sys
import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, S, T, out_dir):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from radae_b200 import rdw, multigpu
    from oracle.core import CoreOraclePort, synth_features, pack_enc_input
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blob = open(rdw.default_weights_path(), "rb").read() if rank == 0 else None
    blob = multigpu.broadcast_weights(dist, rank, blob)
    assert hashlib.sha256(blob).hexdigest() == hashlib.sha256(open(rdw.default_weights_path(), "rb").read()).hexdigest()
    tmp = os.path.join(out_dir, f"w{rank}.rdw"); open(tmp, "wb").write(blob)
    lo, hi = multigpu.stream_shard(S, world, rank)
    x = pack_enc_input(synth_features(S, 4 * T, seed=77))[lo:hi]          # every rank derives its own shard's inputs
    core = CoreOraclePort(rdw_path=tmp, n_streams=hi - lo)
    z = core.encode(x); f = core.decode(z)
    np.save(os.path.join(out_dir, f"z{rank}.npy"), z); np.save(os.path.join(out_dir, f"f{rank}.npy"), f)
    t = multigpu.reduce_max(dist, 1.0 + rank)
    assert t == float(world)
    dist.barrier(); dist.destroy_process_group()


def test_sharded_equals_unsharded(tmp_path):
    import torch.multiprocessing as mp
    from oracle.core import CoreOraclePort, synth_features, pack_enc_input
    from radae_b200 import multigpu
    S, T, world = 7, 5, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, S, T, str(tmp_path)), nprocs=world, join=True)
    x = pack_enc_input(synth_features(S, 4 * T, seed=77))
    core = CoreOraclePort(n_streams=S)
    z = core.encode(x); f = core.decode(z)
    zs = np.concatenate([np.load(tmp_path / f"z{r}.npy") for r in range(world)])
    fs = np.concatenate([np.load(tmp_path / f"f{r}.npy") for r in range(world)])
    assert np.array_equal(z, zs) and np.array_equal(f, fs)
    # shards tile the stream range exactly
    edges = [multigpu.stream_shard(S, world, r) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == S and all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
