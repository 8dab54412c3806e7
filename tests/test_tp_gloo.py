"""Tensor-parallel host logic on CPU: world_size = 2 over gloo.  Covers the split tables, the column shard
of a checkpoint tensor dict (EXL2 and GPTQ) against the oracle, and the gather layouts for rows == 1 (in place) and rows > 1."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_q):
    try:
        for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import exl2_oracle as oracle
        import synth
        from exllamav2_b200.linear import tp_column_slice
        from exllamav2_b200.model import PRESETS
        from exllamav2_b200.tensor_p import TPContext, split_even, tp_column_slice_t

        cfg = PRESETS["test-small"]()
        tp = TPContext(cfg, rank, world)
        assert tp.q == [(0, 256), (256, 512)] and tp.kv == tp.q and tp.rs == [(0, 256), (256, 512)]
        assert tp.id == [(0, 704), (704, 1408)] and tp.vc == [(0, 256), (256, 512)]
        with pytest.raises(ValueError):
            split_even(100, 3, 8)

        # column shard of an EXL2 and a GPTQ matrix: shard-reconstruct == columns of the full reconstruct, bit for bit
        K, N = 256, 512
        a, b = tp.mine(split_even(N, world, 8))
        rng = np.random.default_rng(3)
        x1 = rng.normal(0, 1, size=(1, K)).astype(np.float16)
        x3 = rng.normal(0, 1, size=(3, K)).astype(np.float16)
        for kind in ("exl2", "gptq"):
            if kind == "exl2":
                w = synth.make_exl2(K, N, (5, 4), (0.2, 0.8), 64, seed=11)
                W = oracle.exl2_reconstruct(w)
                Ws = oracle.exl2_reconstruct(tp_column_slice(w, a, b))
            else:
                w = synth.make_gptq(K, N, 128, seed=12, act_order=True)
                W = oracle.gptq_reconstruct(w)
                Ws = oracle.gptq_reconstruct(tp_column_slice(w, a, b))
            assert np.array_equal(W[:, a:b].view(np.uint16), Ws.view(np.uint16)), kind
            # torch slicer == numpy slicer
            wt = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in w.items() if isinstance(v, np.ndarray)}
            st = tp_column_slice_t(wt, a, b)
            sn = tp_column_slice(w, a, b)
            for k in st:
                assert np.array_equal(st[k].numpy(), sn[k]), (kind, k)
            # sharded GEMM + all-gather == unsharded GEMM (no all-reduce: each output element has its single-GPU summation)
            for x in (x1, x3):
                y_full = oracle.gemm_truth(x, W).astype(np.float16)
                y_loc = torch.from_numpy(oracle.gemm_truth(x, Ws).astype(np.float16))
                out = torch.zeros((x.shape[0], N), dtype=torch.half)
                if x.shape[0] == 1:
                    out[:, a:b] = y_loc                 # in place: my slice of the replicated buffer
                    tp.all_gather_cols(out, out[:, a:b])
                else:
                    tp.all_gather_cols(out, y_loc)
                assert np.array_equal(out.numpy().view(np.uint16), y_full.view(np.uint16)), (kind, x.shape)
        dist.barrier()
        dist.destroy_process_group()
        out_q.put((rank, "ok"))
    except Exception as e:          # noqa: BLE001 -- report to the parent
        import traceback
        out_q.put((rank, "".join(traceback.format_exception(type(e), e, e.__traceback__))))


def test_tp_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, res in results:
        assert res == "ok", f"rank {rank}:\n{res}"
