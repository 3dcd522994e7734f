"""surely_raytracing_b200 -- B200 (sm_100a) backend for the per-pixel integration hot path of
carlosconley/surely-raytracing (reference src/render.rs:144-312).  The product is
librtb200.so (include/rtb200.h); this package is the Python plumbing around it."""
from .capi import (PIPELINE_DEFAULT, PIPELINE_MEGAKERNEL, PIPELINE_WAVEFRONT, RTB_FLAG_ISO_PDF_ZERO,
                   RTB_FLAG_PROPAGATE_NAN, RTB_TRACE_BRUTE_FORCE, VARIANT_LIGHTS, RtbError, load_library)
from .api import Scene, auto_expose, render_multi
from .scenes import CONFIGS, DEFAULT_SEED, BuiltScene

__all__ = ["Scene", "render_multi", "auto_expose", "BuiltScene", "CONFIGS", "DEFAULT_SEED", "RtbError", "load_library",
           "PIPELINE_DEFAULT", "PIPELINE_MEGAKERNEL", "PIPELINE_WAVEFRONT", "RTB_FLAG_ISO_PDF_ZERO",
           "RTB_FLAG_PROPAGATE_NAN", "RTB_TRACE_BRUTE_FORCE", "VARIANT_LIGHTS"]
