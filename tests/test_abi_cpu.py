"""CPU: the C-ABI library loads and exports every symbol include/*.h declares;
record layouts agree between C, ctypes and numpy; repo layout rules hold."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "traceml_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tml_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from traceml_b200 import _abi

    lib = _abi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in traceml_b200.h but not exported"
        assert name in _abi.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.tml_abi_version() == 1
    assert lib.tml_status_str(-5) == b"step ids decrease inside the ring"


def test_exported_symbols_are_unmangled_c():
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "traceml_b200", "libtraceml_b200.so")],
                         capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    for name in _declared_symbols():
        assert name in exported


def test_record_layouts_agree():
    from traceml_b200 import _abi
    from traceml_b200.records import PROC_RECORD_DTYPE, STEP_RECORD_DTYPE, WINDOW_ROW_DTYPE

    assert C.sizeof(_abi.StepRecord) == STEP_RECORD_DTYPE.itemsize == 128
    assert C.sizeof(_abi.ProcRecord) == PROC_RECORD_DTYPE.itemsize == 64
    assert WINDOW_ROW_DTYPE.itemsize == 64
    for f, _ in _abi.StepRecord._fields_:
        assert getattr(_abi.StepRecord, f).offset == STEP_RECORD_DTYPE.fields[f][1], f
    for f, _ in _abi.ProcRecord._fields_:
        assert getattr(_abi.ProcRecord, f).offset == PROC_RECORD_DTYPE.fields[f][1], f
    hdr = open(os.path.join(ROOT, "include", "traceml_b200.h")).read()
    assert "128 B / step / rank" in hdr and "64 B / sample / rank" in hdr


def test_null_and_bad_arguments_are_rejected_without_a_gpu():
    from traceml_b200 import _abi

    lib = _abi.lib()
    assert lib.tml_phase_host(None, 0, 1) == -2
    assert lib.tml_phase_begin(None, 0, None) == -2
    assert lib.tml_step_commit(None, 1, 0, 0, 0, 0.0, None) == -2
    assert lib.tml_step_count(None) == 0
    assert lib.tml_shutdown(None) == 0
    buf = C.create_string_buffer(8)
    din = _abi.StDiagIn()
    din.n_ranks = 0
    assert lib.tml_diag_step_time(C.byref(din), buf, 2) == -8  # TML_ERR_SMALL


def test_product_never_imports_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "traceml_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(base, f))
    assert not bad, f"product files import the oracle: {bad}"


def test_no_cpu_fallback_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("needs a box without CUDA")
    from traceml_b200 import runtime
    from traceml_b200.sdk import initial

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        runtime.get_engine()
    initial._reset_for_tests()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        initial.init(mode="manual")
    assert initial.get_init_config() is None


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from traceml_b200 import _abi

    monkeypatch.setattr(_abi, "_LIB", None)
    monkeypatch.setattr(_abi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _abi.lib()


def test_replay_digest_is_stable():
    import replay

    a = replay.make_step_replay("ragged", 3, 50, seed=8)
    b = replay.make_step_replay("ragged", 3, 50, seed=8)
    assert replay.replay_digest(a) == replay.replay_digest(b)
    assert replay.replay_digest(a) != replay.replay_digest(replay.make_step_replay("ragged", 3, 50, seed=9))
    only = replay.make_step_replay("ragged", 3, 50, seed=8, only_ranks=[2])
    assert np.array_equal(only[2], a[2])


def test_ctypes_mirrors_match_the_c_layouts():
    """Every struct that crosses the ABI: sizeof in the library == sizeof of the ctypes mirror."""
    from traceml_b200 import _abi

    lib = _abi.lib()
    pairs = {
        "tml_step_record": _abi.StepRecord, "tml_proc_record": _abi.ProcRecord,
        "tml_live_phase": _abi.LivePhase, "tml_live_stats": _abi.LiveStats,
        "tml_win_info": _abi.WinInfo, "tml_align_info": _abi.AlignInfo,
        "tml_reduce_args": _abi.ReduceArgs, "tml_band_args": _abi.BandArgs, "tml_band_out": _abi.BandOut,
        "tml_proc_agg": _abi.ProcAgg, "tml_comm": _abi.Comm, "tml_reduce_run_args": _abi.ReduceRunArgs,
        "tml_kind_result": _abi.KindResultC, "tml_reduce_run_out": _abi.ReduceRunOut,
        "tml_combined_info": _abi.CombinedInfo, "tml_combined_align": _abi.CombinedAlign,
        "tml_rank_means": _abi.RankMeans, "tml_trend_in": _abi.TrendIn, "tml_st_diag_in": _abi.StDiagIn,
        "tml_mem_metric_in": _abi.MemMetricIn, "tml_mem_diag_in": _abi.MemDiagIn,
        "tml_proc_diag_in": _abi.ProcDiagIn, "tml_sections_args": _abi.SectionsArgs,
        "tml_layer_record": _abi.LayerRecord,
    }
    for name, cls in pairs.items():
        assert int(lib.tml_struct_size(name.encode())) == C.sizeof(cls), name
    assert int(lib.tml_struct_size(b"nope")) == 0
