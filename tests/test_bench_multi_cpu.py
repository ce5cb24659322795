"""The N > 1 aggregation of bench.py on CPU: world_size 2, gloo.  Streams shard into disjoint slices, the
job time is the MAX over ranks, and throughput counts every rank's frames."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, last = bench.stream_range(rank, 256)
        owned = torch.zeros(world * 256, dtype=torch.int32)
        owned[first:last] = 1
        dist.all_reduce(owned)                     # every stream owned exactly once across ranks
        ms = bench.max_over_ranks(10.0 + 5.0 * rank, world, "cpu")
        fps = bench.job_throughput(world, 256, 30, ms)
        out.put((rank, bool((owned == 1).all()), ms, fps))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_max_time():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, disjoint, ms, fps in res:
        assert disjoint
        assert ms == 15.0                          # the slower rank's time on every rank
        assert abs(fps - 2 * 256 * 30 / 15e-3) < 1e-6


def test_reference_arm_other_ranks_exit_quietly(monkeypatch, capsys):
    import bench
    monkeypatch.setenv("RANK", "1")
    args = type("A", (), dict(ref_streams=1, steps=1, warmup=1, workload="mimi", gpus=2, streams=256, prefill=0))()
    bench.run_reference(args)
    assert capsys.readouterr().out == ""


def test_step_algorithmic_bytes_match_survey_8d():
    """stt-1b, 256 streams, bf16 ring: LM 26.88 GB + Mimi 2.36 GB per step."""
    import bench
    from dsm_b200.spec import stt_1b_en_fr
    spec = stt_1b_en_fr()
    B = 256
    kv = 2 * B * spec.num_layers * spec.num_heads * spec.context * spec.head_dim * 2      # K and V, bf16
    total, parts = bench.step_algorithmic_bytes("asr", B, spec, kv)
    lm = sum(v for k, v in parts.items() if k.startswith("lm_"))
    mimi = sum(v for k, v in parts.items() if k.startswith("mimi_"))
    assert abs(bench.lm_streamed_weight_bytes(spec) - 1.677e9) < 2e6
    assert abs(lm - 26.88e9) < 0.05e9 and abs(mimi - 2.36e9) < 0.03e9 and abs(total - lm - mimi) < 1
    assert bench.percentile(list(range(101)), 0.99) == 99 and bench.percentile([5.0], 0.5) == 5.0
