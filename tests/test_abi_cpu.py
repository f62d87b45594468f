"""CPU suite: the C-ABI library loads, exports every symbol include/vr.h declares, and its host-only logic works.
No compute call is made (there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

from cl_volume_renderer_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vr_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 40
    l = C.CDLL(api.LIB_PATH)
    for n in names:
        assert hasattr(l, n), f"libvr.so does not export {n}"
        assert n in api.SYMBOLS, f"api.py does not bind {n}"
    assert sorted(api.SYMBOLS) == names


def test_tf_parse_ui_generated_text():
    # text exactly as app/ui.cpp:160-168 + app/tf_part.cpp:55-79 print it
    src = ("inline bool is_event_gen(short value, short gradient, int4 *color){\n"
           "  if(value >= 500 && value <= 1200)\n {\n    int4 tmp_color = {255,255,255,255};\n"
           "    *color = tmp_color;\n    return true;\n }\n"
           "  if(value >= 812.5 && value <= 3000 && gradient > 12.5 && gradient < 4000)\n {\n"
           "    int4 tmp_color = {25,0,127,76};\n    *color = tmp_color;\n    return true;\n }\n"
           "  \n  return false;\n}\n")
    r = api.tf_parse(src)
    assert len(r) == 2
    assert (r[0]["min_v"], r[0]["max_v"], r[0]["flags"], r[0]["rgba"]) == (500.0, 1200.0, 0, (255, 255, 255, 255))
    assert (r[1]["min_v"], r[1]["min_g"], r[1]["max_g"], r[1]["flags"]) == (812.5, 12.5, 4000.0, api.VR_TF_USE_GRADIENT)
    assert r[1]["rgba"] == (25, 0, 127, 76)
    assert api.tf_parse(api.tf_format(r)) == r  # round trip


def test_tf_parse_threshold_form():
    # tests/sdf/sdf_test.cpp:22 / app/sdf_benchmark.cpp:18
    r = api.tf_parse("inline bool is_event_gen(short value, short gradient, uint4 *color){ return (value > 800); }")
    assert len(r) == 1 and r[0]["flags"] == api.VR_TF_THRESHOLD and r[0]["min_v"] == 800.0


def test_tf_parse_empty_and_rejects():
    assert api.tf_parse("inline bool is_event_gen(short value, short gradient, int4 *color){\n  \n  return false;\n}\n") == []
    for bad in ["", "bool f(){}", "inline bool is_event_gen(short value, short gradient, int4 *color){ return true; }",
                "inline bool is_event_gen(short value, short gradient, int4 *color){ if(value >= 1) {} return false; }"]:
        with pytest.raises(api.VrError):
            api.tf_parse(bad)


def test_no_gpu_fails_loudly_without_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(api.VrError, match="no CPU fallback"):
        api.Context(0)
