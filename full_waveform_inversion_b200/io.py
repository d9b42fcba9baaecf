"""Callers and data formats either side of the Track A hot path and the ``run`` glue.

Same names and signatures as the reference module (cited FWI:<line>).  File parsing stays on the host (NumPy);
the Green's-function conditioning - time shift with zeroed head, phase-window cut, unit scaling - is one device op
(``fwi_mc_prepare``, float64, bit-identical to the NumPy operations it replaces).  Output dictionaries keep the
reference's keys so ``plot_full_waveform_inversion.py`` (PLOT:67-77) and the UNP script (UNP:53-58) read them.
"""
from __future__ import annotations

import math
import os
import pickle
from ctypes import c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import check, current_stream, ptr
from . import full_waveform_inversion as fw

_COMBINED = ("DC_single_force_couple", "DC_single_force_no_coupling", "DC_crack_couple", "single_force_crack_no_coupling")


# ------------------------------------------------------------------------------------------------ f2: inputs
def prepare_green_functions(raw, manual_indices_time_shift=(), cut_phase_start_vals=(), cut_phase_length=0,
                            set_pre_time_shift_values_to_zero_switch=True, scale1=1.0, scale2=1.0):
    """Device op behind the loaders: roll + zero head (FWI:94-101), cut (FWI:104-111), scale (FWI:178-196).
    raw: float64 (K,C,T) or (K,C,T,2) -> float64 array of the same rank."""
    lib = _lib.require_gpu()
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    K, C, T = raw.shape[:3]
    nm = 1 if raw.ndim == 3 else raw.shape[3]
    dev = torch.device("cuda", torch.cuda.current_device())
    raw_d = torch.from_numpy(raw).to(dev)
    shift_d = cut_d = None
    if len(manual_indices_time_shift) > 0:
        if len(manual_indices_time_shift) > K:
            raise IndexError("manual_indices_time_shift has %d entries for %d traces (FWI:97 indexes past the array)" % (len(manual_indices_time_shift), K))
        sh = np.full(K, np.iinfo(np.int32).min, dtype=np.int32)     # sentinel: trace left zero, as FWI:95-96 does for missing shifts
        sh[: len(manual_indices_time_shift)] = np.asarray(manual_indices_time_shift, dtype=np.int32)
        shift_d = torch.from_numpy(sh).to(dev)
    Tout = T
    if len(cut_phase_start_vals) > 0:
        if len(cut_phase_start_vals) != K or cut_phase_length < 1:
            raise ValueError("cut_phase_start_vals needs one start per trace and cut_phase_length >= 1")
        cut_d = torch.from_numpy(np.asarray(cut_phase_start_vals, dtype=np.int64).astype(np.int32)).to(dev)
        Tout = int(cut_phase_length)
    out = torch.empty((K, C, Tout) + ((nm,) if raw.ndim == 4 else ()), dtype=torch.float64, device=dev)
    check(lib.fwi_mc_prepare(ptr(raw_d), K, C, T, nm, ptr(shift_d), 1 if set_pre_time_shift_values_to_zero_switch else 0,
                             ptr(cut_d), Tout, float(scale1), float(scale2), ptr(out), current_stream()))
    return out.cpu().numpy()


def _read_traces(datadir, real_data_fnames):
    return np.stack([np.loadtxt(datadir + "/" + f, dtype=float) for f in real_data_fnames])          # FWI:88-89


def _cut_data(real_data_array, cut_phase_start_vals, cut_phase_length):
    if len(cut_phase_start_vals) == 0:
        return real_data_array
    return np.stack([real_data_array[i, int(s): int(s) + int(cut_phase_length)] for i, s in enumerate(cut_phase_start_vals)])   # FWI:108


def load_input_data(datadir, real_data_fnames, green_func_fnames, manual_indices_time_shift=[], cut_phase_start_vals=[],
                    cut_phase_length=0, set_pre_time_shift_values_to_zero_switch=True, _scale=(1.0, 1.0)):
    """Load data traces and Green's functions, shift / cut the Green's functions (FWI:75-113) -> (real (K,T), G (K,C,T))."""
    real = _read_traces(datadir, real_data_fnames)
    raw = np.stack([np.transpose(np.loadtxt(datadir + "/" + f, dtype=float)) for f in green_func_fnames])   # FWI:90
    G = prepare_green_functions(raw, manual_indices_time_shift, cut_phase_start_vals, cut_phase_length,
                                set_pre_time_shift_values_to_zero_switch, *_scale)
    return _cut_data(real, cut_phase_start_vals, cut_phase_length), G


def load_input_data_multiple_media(datadir, real_data_fnames, green_func_fnames, green_func_fnames_split_index,
                                   manual_indices_time_shift=[], cut_phase_start_vals=[], cut_phase_length=0,
                                   set_pre_time_shift_values_to_zero_switch=True, _scale=(1.0, 1.0)):
    """Two-media variant (FWI:116-165): Green's functions gain a trailing axis of size 2."""
    f1 = green_func_fnames[:green_func_fnames_split_index]
    f2 = green_func_fnames[green_func_fnames_split_index:]
    if len(f1) != len(f2):
        raise ValueError("Greens functions fname array is not correct. Consider whether green_func_fnames_split_index "
                         "value is correct for splitting the two mediums.")                                # FWI:123-125
    real = _read_traces(datadir, real_data_fnames)
    raw = np.stack([np.stack([np.transpose(np.loadtxt(datadir + "/" + a, dtype=float)),
                              np.transpose(np.loadtxt(datadir + "/" + b, dtype=float))], axis=-1) for a, b in zip(f1, f2)])
    G = prepare_green_functions(raw, manual_indices_time_shift, cut_phase_start_vals, cut_phase_length,
                                set_pre_time_shift_values_to_zero_switch, *_scale)
    return _cut_data(real, cut_phase_start_vals, cut_phase_length), G


def get_overall_real_and_green_func_data(datadir, real_data_fnames, MT_green_func_fnames, single_force_green_func_fnames,
                                         inversion_type, manual_indices_time_shift_MT=[], manual_indices_time_shift_SF=[],
                                         cut_phase_start_vals=[], cut_phase_length=0,
                                         set_pre_time_shift_values_to_zero_switch=True,
                                         invert_for_ratio_of_multiple_media_greens_func_switch=False,
                                         green_func_fnames_split_index=0):
    """Per inversion type: which Green's functions to load, MT x 1e3, MT (+) SF stacking, x 1e7 (FWI:168-197)."""
    kw = dict(cut_phase_start_vals=cut_phase_start_vals, cut_phase_length=cut_phase_length,
              set_pre_time_shift_values_to_zero_switch=set_pre_time_shift_values_to_zero_switch)
    multi = invert_for_ratio_of_multiple_media_greens_func_switch

    def load(fnames, shifts, scale):
        if multi:
            return load_input_data_multiple_media(datadir, real_data_fnames, fnames, green_func_fnames_split_index, shifts,
                                                  _scale=scale, **kw)
        return load_input_data(datadir, real_data_fnames, fnames, shifts, _scale=scale, **kw)

    if inversion_type in ("full_mt", "DC", "DC_crack_couple"):
        return load(MT_green_func_fnames, manual_indices_time_shift_MT, (1e3, 1e7))                       # FWI:178, FWI:196
    if inversion_type == "single_force":
        return load(single_force_green_func_fnames, manual_indices_time_shift_SF, (1.0, 1e7))             # FWI:196
    if inversion_type in ("DC_single_force_couple", "DC_single_force_no_coupling", "single_force_crack_no_coupling"):
        real, G_mt = load(MT_green_func_fnames, manual_indices_time_shift_MT, (1e3, 1e7))                 # FWI:192, 196
        _, G_sf = load(single_force_green_func_fnames, manual_indices_time_shift_SF, (1.0, 1e7))
        return real, np.hstack((G_mt, G_sf))                                                               # FWI:194
    raise ValueError("unknown inversion_type %r" % (inversion_type,))


# ------------------------------------------------------------------------------------------------ f3: outputs
def get_event_uid_and_station_data_MTFIT_FORMAT_from_nonlinloc_hyp_file(nlloc_hyp_filename):
    """uid and MTFIT-style station list from a NonLinLoc .hyp file (FWI:872-946), parsed in Python instead of
    shelling out to grep / awk; the origin time is formatted like obspy's strftime('%Y%m%d%H%M%S%f')."""
    geo, phases, in_phase = None, [], False
    with open(nlloc_hyp_filename) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "GEOGRAPHIC":
                geo = tok
            if tok[0] == "END_PHASE":
                in_phase = False
            if in_phase:
                phases.append(tok)
            if tok[0] == "PHASE" and len(tok) > 1 and tok[1] == "ID":
                in_phase = True
    if geo is None:
        raise ValueError("no GEOGRAPHIC line in %s" % nlloc_hyp_filename)
    sec = float(geo[7])
    whole = int(math.floor(sec))
    micro = int(round((sec - whole) * 1e6))
    uid = "%04d%02d%02d%02d%02d%02d%06d" % (int(geo[2]), int(geo[3]), int(geo[4]), int(geo[5]), int(geo[6]), whole, micro)
    angles = {}
    for tok in phases:                                   # first P pick per station defines azimuth / take-off (FWI:922-935)
        if tok[4] == "P":
            angles[tok[0]] = (float(tok[22]), 180.0 - float(tok[24]))
    stations = [[np.array([sta], dtype=str), np.array([[azi]], dtype=float), np.array([[toa]], dtype=float),
                 np.array([[0]], dtype=int)] for sta, (azi, toa) in angles.items()]
    return uid, stations


def remove_zero_prob_results(MTp, MTs):
    """Drop samples with zero probability (FWI:948-953)."""
    keep = np.nonzero(np.asarray(MTp) > 0.0)[0]
    return np.asarray(MTp)[keep], np.asarray(MTs)[:, keep]


def save_to_MTFIT_style_file(MTs, MTp, nlloc_hyp_filename, inversion_type, outdir, MTp_absolute=[]):
    """Pickled dict {MTs, MTp, uid, stations[, MTp_absolute]} -> <outdir>/<uid>_FW_<type>.pkl (FWI:955-971)."""
    uid, stations = get_event_uid_and_station_data_MTFIT_FORMAT_from_nonlinloc_hyp_file(nlloc_hyp_filename)
    out = {"MTs": MTs, "MTp": MTp, "uid": uid, "stations": stations}
    if len(MTp_absolute) > 0:
        out["MTp_absolute"] = MTp_absolute
    fname = outdir + "/" + uid + "_FW_" + inversion_type + ".pkl"
    with open(fname, "wb") as fh:
        pickle.dump(out, fh)
    return fname


def save_specific_waveforms_to_file(real_data_array, synth_data_array, data_labels, nlloc_hyp_filename, inversion_type, outdir):
    """Pickled dict label -> {real_wf, synth_wf} -> <outdir>/<uid>_FW_<type>.wfs (FWI:1022-1035)."""
    out = {lab: {"real_wf": real_data_array[i, :], "synth_wf": synth_data_array[i, :]} for i, lab in enumerate(data_labels)}
    uid, _ = get_event_uid_and_station_data_MTFIT_FORMAT_from_nonlinloc_hyp_file(nlloc_hyp_filename)
    fname = outdir + "/" + uid + "_FW_" + inversion_type + ".wfs"
    with open(fname, "wb") as fh:
        pickle.dump(out, fh)
    return fname


# ------------------------------------------------------------------------------------------------ run glue
def _n_phase_types(labels):
    return sum(1 for k in ("P", "S", "surface") if list(labels).count(k) > 0)                               # FWI:1050-1056


def run(datadir, outdir, real_data_fnames, MT_green_func_fnames, single_force_green_func_fnames, data_labels, inversion_type,
        perform_normallised_waveform_inversion, compare_all_waveforms_simultaneously, num_samples, comparison_metric,
        manual_indices_time_shift_MT, manual_indices_time_shift_SF, nlloc_hyp_filename, cut_phase_start_vals=[],
        cut_phase_length=0, plot_switch=False, num_processors=1, set_pre_time_shift_values_to_zero_switch=True,
        only_save_non_zero_solns_switch=False, return_absolute_similarity_values_switch=False,
        invert_for_ratio_of_multiple_media_greens_func_switch=False, green_func_fnames_split_index=0,
        green_func_phase_labels=[], seed=0):
    """The reference's `run` / `run_multi_medium_inversion` (FWI:1161-1232, FWI:1037-1158): load -> least squares ->
    save LSQ result -> Monte-Carlo sampling -> save.  Plotting is out of scope (`plot_switch` is ignored).
    Returns (MTs, MTp, MTp_absolute) in addition to writing the reference's files."""
    multi = invert_for_ratio_of_multiple_media_greens_func_switch
    real, G = get_overall_real_and_green_func_data(
        datadir, real_data_fnames, MT_green_func_fnames, single_force_green_func_fnames, inversion_type,
        manual_indices_time_shift_MT=manual_indices_time_shift_MT, manual_indices_time_shift_SF=manual_indices_time_shift_SF,
        cut_phase_start_vals=cut_phase_start_vals, cut_phase_length=cut_phase_length,
        set_pre_time_shift_values_to_zero_switch=set_pre_time_shift_values_to_zero_switch,
        invert_for_ratio_of_multiple_media_greens_func_switch=multi, green_func_fnames_split_index=green_func_fnames_split_index)
    n_phase = 0
    if multi:
        if len(green_func_phase_labels) > 0 and len(green_func_phase_labels) != G.shape[0]:
            raise ValueError("Greens functions filename array (for medium 1), does not match length of green_func_phase_labels array.")
        n_phase = _n_phase_types(green_func_phase_labels)
        G_lsq = 0.5 * G[..., 0] + 0.5 * G[..., 1]                                                            # FWI:1059-1060
    else:
        G_lsq = G
    M = fw.perform_inversion(real, G_lsq)                                                                    # FWI:1175
    M_amplitude = float(np.sum(M ** 2) ** 0.5)                                                               # FWI:1176
    synth = fw.forward_model(G_lsq, M)
    sim = fw.compare_synth_to_real_waveforms(real, synth, comparison_metric, perform_normallised_waveform_inversion,
                                             compare_all_waveforms_simultaneously)
    lsq_dir = outdir + "/least_squares_result"
    os.makedirs(lsq_dir, exist_ok=True)
    save_to_MTFIT_style_file(M, np.array([sim]), nlloc_hyp_filename, inversion_type, lsq_dir)                # FWI:1191-1193
    save_specific_waveforms_to_file(real, synth, data_labels, nlloc_hyp_filename, inversion_type, lsq_dir)   # (q5 not reproduced: all components)

    MTs, MTp, MTp_abs = fw.perform_monte_carlo_sampled_waveform_inversion(
        real, G, num_samples, M_amplitude=M_amplitude, inversion_type=inversion_type, comparison_metric=comparison_metric,
        perform_normallised_waveform_inversion=perform_normallised_waveform_inversion,
        compare_all_waveforms_simultaneously=compare_all_waveforms_simultaneously, num_processors=num_processors,
        return_absolute_similarity_values_switch=return_absolute_similarity_values_switch,
        invert_for_ratio_of_multiple_media_greens_func_switch=multi, green_func_phase_labels=green_func_phase_labels,
        num_phase_types_for_media_ratios=n_phase, seed=seed)                                                 # FWI:1203 / FWI:1091
    if only_save_non_zero_solns_switch:
        MTp, MTs = remove_zero_prob_results(MTp, MTs)                                                        # FWI:1211-1212
    os.makedirs(outdir, exist_ok=True)
    save_to_MTFIT_style_file(MTs, MTp, nlloc_hyp_filename, inversion_type, outdir, MTp_absolute=MTp_abs)
    best = fw.get_synth_forward_model_most_likely_result(MTs, MTp, G, inversion_type, multi, green_func_phase_labels, n_phase)
    save_specific_waveforms_to_file(real, best, data_labels, nlloc_hyp_filename, inversion_type, outdir)
    return MTs, MTp, MTp_abs


def run_multi_medium_inversion(*args, **kwargs):
    """FWI:1037-1158 - same argument list as `run`, with the two-media switch forced on."""
    kwargs["invert_for_ratio_of_multiple_media_greens_func_switch"] = True
    return run(*args, **kwargs)


def load_MT_dict_from_file(filename):
    """Read back a .pkl written by save_to_MTFIT_style_file (UNP:52-58 / PLOT:67-77) -> (uid, MTp, MTs, stations)."""
    with open(filename, "rb") as fh:
        d = pickle.load(fh)
    return d["uid"], d["MTp"], d["MTs"], d["stations"]


def unnormallised_probability_run(inversion_type, event_uid, datadir_FW_outputs, datadir_greens_functions, real_data_fnames,
                                  MT_green_func_fnames, single_force_green_func_fnames, manual_indices_time_shift,
                                  comparison_metric, perform_normallised_waveform_inversion,
                                  compare_all_waveforms_simultaneously):
    """A working version of the UNP script's `run` (UNP:234-264; the original has a duplicated parameter name and
    cannot compile, q7): similarity of the most likely saved sample."""
    real, G = get_overall_real_and_green_func_data(datadir_greens_functions, real_data_fnames, MT_green_func_fnames,
                                                   single_force_green_func_fnames, inversion_type, manual_indices_time_shift,
                                                   manual_indices_time_shift)
    uid, MTp, MTs, _ = load_MT_dict_from_file(datadir_FW_outputs + "/" + event_uid + "_FW_" + inversion_type + ".pkl")
    best = np.asarray(MTs)[:, int(np.argmax(MTp))]
    if inversion_type in _COMBINED:
        best = best[:-1]                                                                                     # UNP:253-254 (extended to every combined type)
    return fw.get_unnormallised_prob_for_specific_soln(real, G, best[: G.shape[1]], comparison_metric,
                                                       perform_normallised_waveform_inversion,
                                                       compare_all_waveforms_simultaneously)
