"""B200-native all-pairs genome *selection* (CB -> smh_a / hll_a / hll_an -> HLL union -> Jaccard >= tau).

Drop-in for the hot path of sanhue903/CUDA_Selection_Criteria (src/selection.cpp:241-291).
The compute lives in the in-tree CUDA library ``libselb200.so`` (csrc/, sm_100a) behind the C-ABI
of ``include/selb200.h``; this package is the thin host mirror of the reference's CLI flow.
There is no CPU fallback: importing works anywhere, computing needs a GPU.
"""
from ._lib import LibraryNotBuilt, SelB200Error, lib, lib_path  # noqa: F401
from .selection import (  # noqa: F401
    CRITERIA,
    Selection,
    SelectionResult,
    band_params,
    format_lines,
    run_filelist,
)
from . import build_sketch, sketch_io, synth  # noqa: F401

__all__ = [
    "Selection", "SelectionResult", "CRITERIA", "band_params", "format_lines", "run_filelist",
    "sketch_io", "synth", "build_sketch", "lib", "lib_path", "SelB200Error", "LibraryNotBuilt",
]
