"""CPU: the mbarrier protocol of the fused split kernels (mlp_fused_split.cu) replayed in a discrete-event model
(tests/protocol_model.py: producer, MMA issuer and the eight epilogue warps as coroutines with the kernels'
own parity expressions, a FIFO tensor pipe, hazard checks on tensor memory).  The shipped "two-instalment" schedule must be
deadlock- and hazard-free over several tiles for the forward (skip layer, condition layer) and the dgrad chain, for every
ring depth the kernels use, and the checker must notice a broken protocol."""
import importlib.util
from pathlib import Path

import pytest

SIM = Path(__file__).resolve().parent / "protocol_model.py"


def _load(text=None):
    if text is None:
        spec = importlib.util.spec_from_file_location("protocol_model", SIM)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod.__dict__
    ns = {}
    exec(compile(text, "sim_mutant", "exec"), ns)
    return ns


@pytest.mark.parametrize("kind", ["forward", "dgrad"])
@pytest.mark.parametrize("NS", [5, 6])
def test_shipped_schedule_is_deadlock_and_hazard_free(kind, NS):
    m = _load()
    total, per_tile, util = m["run"](kind, False, n_tiles=5, NS=NS)
    n_layers = len(m["network"](kind))
    # 96 MMAs of 73 cycles per trunk layer = 7.0 k: the schedule may not be faster than the tensor pipe nor absurdly slower
    assert 6500 <= per_tile / n_layers <= 9000
    assert 0.7 <= util <= 1.0


@pytest.mark.parametrize("timing", [(20, 50, 400), (230, 4000, 400), (230, 1220, 2500), (5, 5, 5)])
def test_shipped_schedule_under_other_timings(timing):
    m = _load()
    m["T_LD"], m["T_CHUNK"], m["T_TMA"] = map(float, timing)
    for kind in ("forward", "dgrad"):
        m["run"](kind, False, n_tiles=4, NS=5)


def test_the_model_notices_a_broken_protocol():
    src = SIM.read_text()
    # the issuer of the shipped kernel skipping its act_ready wait before k-block 2 must read a stale ACT
    bad = src.replace('                    if wait_act and kb == st.n_act_kb // 2:\n                        yield ("wait", S.act_ready, n_act & 1)\n', "")
    assert bad != src
    m = _load(bad)
    m["T_LD"], m["T_CHUNK"] = 230.0, 4000.0
    with pytest.raises(m["Hazard"]):
        m["run"]("forward", False, n_tiles=3, NS=5)
    # a flipped parity on the accumulator hand-back deadlocks
    bad = src.replace('yield ("wait", S.acc_empty[h], (n_acc[h] & 1) ^ 1)', 'yield ("wait", S.acc_empty[h], n_acc[h] & 1)')
    assert bad != src
    m = _load(bad)
    with pytest.raises(m["Deadlock"]):
        m["run"]("forward", False, n_tiles=2, NS=5)
