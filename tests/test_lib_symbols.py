"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol that include/fsw_embedding.h declares (and the ctypes table binds each of them)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "fsw_embedding.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|size_t|int64_t|char\*|const char\*)\s+\*?\s*([a-z_0-9]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_reference_abi():
    names = header_functions()
    # the 7 symbols of the reference's libfsw_embedding.so (fsw_embedding.cu:125-183, :194, :212, :231)
    for legacy in ["segcumsum_wrapper", "add_block_sums_wrapper", "get_max_threads_per_block",
                   "launch_segcumsum_kernel_float", "launch_segcumsum_kernel_double",
                   "launch_add_block_sums_kernel_float", "launch_add_block_sums_kernel_double"]:
        assert legacy in names
    for new in ["fsw_embed_forward", "fsw_embed_backward", "fsw_gemm", "fsw_segcumsum", "fsw_csr_from_edge_index",
                "fsw_segment_plan", "fsw_last_error"]:
        assert new in names


def test_library_loads_and_exports_every_declared_symbol():
    from fsw_gnn_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -m fsw_gnn_b200.build"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), "libfsw_embedding.so does not export %s" % name
    bound = _lib.load()
    assert set(header_functions()) == set(_lib.EXPORTED_SYMBOLS)
    assert bound.fsw_version() >= 100
    assert bound.fsw_built_for_sm() == 100
    assert bound.fsw_last_error() is not None


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing elsewhere."""
    import pytest
    import torch
    from fsw_gnn_b200 import FSW_embedding, segcumsum
    emb = FSW_embedding(3, 8, device="cpu")
    with pytest.raises(RuntimeError):
        emb(torch.randn(5, 3))
    with pytest.raises(RuntimeError):
        segcumsum(torch.randn(4), torch.zeros(4, dtype=torch.int64))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fsw_gnn_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("no oracle", ""), fn
