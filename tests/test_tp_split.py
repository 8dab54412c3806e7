"""Host-side split tables of the column shard (exllamav2_b200/tensor_p.py TPContext) for every world size the scaling run uses:
each rank's column range of every sharded dimension is a multiple of what the kernels / the checkpoint format need (whole heads,
8 columns per scale word, 32-column blocks), the ranges tile the dimension, and the per-rank matrices keep whole 32-column blocks."""
import pytest

from exllamav2_b200.model import PRESETS
from exllamav2_b200.tensor_p import TPContext, split_even


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("preset", ["llama2-7b-4.0bpw", "llama2-7b-gptq-g128-act"])
def test_split_tables(preset, world):
    cfg = PRESETS[preset]()
    hd = cfg.head_dim
    full = {"kv": cfg.num_kv_heads * hd, "q": cfg.num_heads * hd, "id": cfg.intermediate_size, "rs": cfg.hidden_size, "vc": cfg.vocab_size}
    mult = {"kv": hd, "q": hd, "id": 8, "rs": 32, "vc": 32}
    for rank in range(world):
        tp = TPContext(cfg, rank, world)
        for name, n in full.items():
            table = getattr(tp, name)
            assert len(table) == world and table[0][0] == 0 and table[-1][1] == n
            assert all(table[i][1] == table[i + 1][0] for i in range(world - 1)), "ranges tile the dimension"
            a, b = tp.mine(table)
            assert (b - a) % mult[name] == 0 and a % mult[name] == 0
            assert (b - a) % 8 == 0, "a scale word packs 8 columns"
            if name in ("q", "kv", "rs", "vc"):
                assert (b - a) % 32 == 0, "whole 32-column blocks per rank"


def test_split_even_rejects_ragged():
    with pytest.raises(ValueError):
        split_even(100, 3, 8)
