"""B200-native hot paths of Kevin2599/full_waveform_inversion.

``full_waveform_inversion_b200.full_waveform_inversion`` mirrors the reference module's
function names and signatures for the Monte-Carlo source-inversion path (Track A);
``full_waveform_inversion_b200.acoustic`` holds the finite-difference forward / adjoint /
gradient / model-update path BASELINE.json names (Track B, no reference counterpart).
Both call hand-written sm_100a CUDA through the C ABI in include/fwi_b200.h.
"""
__version__ = "0.1.0"
