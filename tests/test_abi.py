"""The C-ABI library loads and exports every symbol include/ctunet_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ctunet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctu_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_entry_points():
    names = _declared()
    for must in ("ctu_umma_gemm", "ctu_in_apply", "ctu_in_stats", "ctu_layernorm", "ctu_attention", "ctu_pwa_fuse",
                 "ctu_blend_accumulate", "ctu_blend_normalize", "ctu_conv_cin1", "ctu_patchify_ln",
                 "ctu_umma_wgrad", "ctu_in_bwd_stats", "ctu_in_bwd_apply", "ctu_layernorm_bwd", "ctu_attention_bwd",
                 "ctu_pwa_fuse_bwd", "ctu_gelu_bwd", "ctu_colsum", "ctu_accumulate", "ctu_pack_weights", "ctu_unpack_grads",
                 "ctu_dice_ce_fwd", "ctu_dice_ce_bwd", "ctu_ensemble_argmax", "ctu_adamw_step",
                 "ctu_set_persistent_sm_limit", "ctu_ffn_fused", "ctu_dice_ce_finalize", "ctu_gather3d", "ctu_cc_filter_largest", "ctu_stats_fold",
                 "ctu_invert_resample", "ctu_invert_ensemble_argmax"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from hybrid_ctunet_b200 import lib
    handle = lib.load()
    for name in _declared():
        assert hasattr(handle, name), f"{name} declared in include/ctunet_b200.h but not exported"
    assert handle.ctu_version().startswith(b"ctunet_b200")
    assert isinstance(lib.launch_count(), int)


def test_ctypes_struct_matches_header_field_order():
    from hybrid_ctunet_b200.lib import GemmDesc
    src = open(os.path.join(ROOT, "include", "ctunet_b200.h")).read()
    start = src.index("typedef struct ctu_gemm_desc {") + len("typedef struct ctu_gemm_desc {")
    body = src[start:src.index("} ctu_gemm_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        m = re.match(r"(const\s+)?(void|float|double|int32_t)\s*\*?\s*(.*)", decl)
        if m and "{" not in decl:
            fields += [f.strip().lstrip("*") for f in m.group(3).split(",") if f.strip()]
    assert fields == [f[0] for f in GemmDesc._fields_]


def test_wgrad_struct_matches_header_field_order():
    from hybrid_ctunet_b200.lib import WgradDesc
    src = open(os.path.join(ROOT, "include", "ctunet_b200.h")).read()
    start = src.index("typedef struct ctu_wgrad_desc {") + len("typedef struct ctu_wgrad_desc {")
    body = re.sub(r"/\*.*?\*/", "", src[start:src.index("} ctu_wgrad_desc;")], flags=re.S)
    fields = []
    for decl in body.split(";"):
        m = re.match(r"(const\s+)?(void|float|double|int32_t)\s*\*?\s*(.*)", decl.strip())
        if m and "{" not in decl:
            fields += [f.strip().lstrip("*") for f in m.group(3).split(",") if f.strip()]
    assert fields == [f[0] for f in WgradDesc._fields_]


def test_invert_geom_struct_matches_header_field_order():
    from hybrid_ctunet_b200.lib import InvertGeom
    src = open(os.path.join(ROOT, "include", "ctunet_b200.h")).read()
    start = src.index("typedef struct ctu_invert_geom {") + len("typedef struct ctu_invert_geom {")
    body = src[start:src.index("} ctu_invert_geom;")]
    fields = re.findall(r"(?:double|int32_t)\s+(\w+)(?:\[\d+\])?;", body)
    assert fields == [f[0] for f in InvertGeom._fields_]
    assert ctypes.sizeof(InvertGeom) == 12 * 8 + 13 * 4 + 4   # 148 bytes + tail padding to the 8-byte alignment


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hybrid_ctunet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)


def test_no_forbidden_batched_memcpy_symbols():
    bad = ("cudaMemcpy" + "BatchAsync", "cudaMemcpy3D" + "BatchAsync", "cuMemcpy" + "BatchAsync", "cuMemcpy3D" + "BatchAsync")
    for dirpath, _, files in os.walk(ROOT):
        if ".git" in dirpath or "gpurun_out" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh", ".c", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not any(b in text for b in bad), f


def test_loss_heads_struct_matches_header_field_order():
    from hybrid_ctunet_b200.lib import LOSS_MAX_HEADS, LossHeads
    src = open(os.path.join(ROOT, "include", "ctunet_b200.h")).read()
    assert f"#define CTU_LOSS_MAX_HEADS {LOSS_MAX_HEADS}" in src
    start = src.index("typedef struct ctu_loss_heads {") + len("typedef struct ctu_loss_heads {")
    body = re.sub(r"/\*.*?\*/", "", src[start:src.index("} ctu_loss_heads;")], flags=re.S)
    fields = []
    for decl in body.split(";"):
        m = re.match(r"(int32_t|int64_t|double)\s+(.*)", decl.strip())
        if m:
            fields += [re.sub(r"\[.*\]", "", f.strip()) for f in m.group(2).split(",") if f.strip()]
    assert fields == [f[0] for f in LossHeads._fields_]
