"""input preparation and output formats.  CPU part: the oracle restatement of the loaders'
arithmetic against the live-reference golden vectors, and the pure-Python .hyp parser / pickle writers.
GPU part (marked): the device op and the drop-in `run()`."""
import os
import pickle

import numpy as np
import pytest

from oracle import mc_oracle as orc


def _write_files(tmp, g):
    names = {"real": [], "mt": [], "sf": [], "mt2": [], "sf2": []}
    K = g["prep_file_real"].shape[0]
    for k in range(K):
        for key in names:
            fn = "%s_%d.txt" % (key, k)
            np.savetxt(os.path.join(tmp, fn), g["prep_file_" + key][k], fmt="%.17e")
            names[key].append(fn)
    return names


CASES = {
    "prep_a": ("full_mt", lambda g: dict(manual_indices_time_shift_MT=list(g["prep_sh_mt"]))),
    "prep_b": ("single_force_crack_no_coupling", lambda g: dict(manual_indices_time_shift_MT=list(g["prep_sh_mt"]),
                                                                manual_indices_time_shift_SF=list(g["prep_sh_sf"]),
                                                                cut_phase_start_vals=list(g["prep_cuts"]), cut_phase_length=30)),
    "prep_c": ("single_force", lambda g: dict(manual_indices_time_shift_SF=list(g["prep_sh_sf"]),
                                              set_pre_time_shift_values_to_zero_switch=False)),
    "prep_d": ("DC", lambda g: dict(manual_indices_time_shift_MT=list(g["prep_sh_mt"]), cut_phase_start_vals=list(g["prep_cuts"]),
                                    cut_phase_length=25, invert_for_ratio_of_multiple_media_greens_func_switch=True,
                                    green_func_fnames_split_index=5)),
    "prep_e": ("DC_single_force_no_coupling", lambda g: dict()),
    # loader quirks (FWI:94-101): negative shifts zero all but the last |shift| samples; missing shifts leave zero traces
    "prep_f": ("full_mt", lambda g: dict(manual_indices_time_shift_MT=[-3, 0, 5, -2, 7])),
    "prep_g": ("single_force", lambda g: dict(manual_indices_time_shift_SF=[3, 1])),
}


def test_oracle_preparation_matches_reference(golden_a):
    g = golden_a
    mt = np.transpose(g["prep_file_mt"], (0, 2, 1))
    sf = np.transpose(g["prep_file_sf"], (0, 2, 1))
    a = orc.prepare_green_functions(mt, g["prep_sh_mt"], scale1=1e3, scale2=1e7)
    assert np.array_equal(a, g["prep_a_G"])
    b = np.hstack((orc.prepare_green_functions(mt, g["prep_sh_mt"], g["prep_cuts"], 30, scale1=1e3, scale2=1e7),
                   orc.prepare_green_functions(sf, g["prep_sh_sf"], g["prep_cuts"], 30, scale2=1e7)))
    assert np.array_equal(b, g["prep_b_G"])
    c = orc.prepare_green_functions(sf, g["prep_sh_sf"], zero_head=False, scale2=1e7)
    assert np.array_equal(c, g["prep_c_G"])
    mt2 = np.stack([mt, np.transpose(g["prep_file_mt2"], (0, 2, 1))], -1)
    d = orc.prepare_green_functions(mt2, g["prep_sh_mt"], g["prep_cuts"], 25, scale1=1e3, scale2=1e7)
    assert np.array_equal(d, g["prep_d_G"])
    f = orc.prepare_green_functions(mt, [-3, 0, 5, -2, 7], scale1=1e3, scale2=1e7)
    assert np.array_equal(f, g["prep_f_G"]) and np.count_nonzero(f[0, :, :-3]) == 0 and np.count_nonzero(f[0, :, -3:]) > 0
    gg = orc.prepare_green_functions(sf, [3, 1], scale2=1e7)
    assert np.array_equal(gg, g["prep_g_G"]) and np.count_nonzero(gg[2:]) == 0


HYP = """NLLOC "loc" "LOCATED" "Location completed."
GEOGRAPHIC  OT 2018 02 14  18 55 38.216400  Lat 46.5 Long 8.3 Depth 0.2
PHASE ID Ins Cmp On Pha  FM Date     HrMn   Sec     Err  ErrMag    Coda      Amp       Per  >   TTpred    Res       Weight    StaLoc(X  Y         Z)        SDist    SAzim  RAz  RDip RQual    Tcorr
RA51   ?    ?    ? P      ? 20180214 1855   38.3000 GAU  2.00e-03 -1.00e+00 -1.00e+00 -1.00e+00 >     0.0831 0.0005    1.0000    1.2000    2.3000   -2.5000    0.1500 123.40 123.4  35.0  9     0.0000
RA51   ?    ?    ? S      ? 20180214 1855   38.4000 GAU  2.00e-03 -1.00e+00 -1.00e+00 -1.00e+00 >     0.1631 0.0005    1.0000    1.2000    2.3000   -2.5000    0.1500 123.40 123.4  35.0  9     0.0000
RA52   ?    ?    ? P      ? 20180214 1855   38.3100 GAU  2.00e-03 -1.00e+00 -1.00e+00 -1.00e+00 >     0.0931 0.0005    1.0000    1.4000    2.1000   -2.5000    0.1700 201.70 201.7  41.5  9     0.0000
END_PHASE
END_NLLOC
"""


def test_hyp_parser_and_writers(tmp_path):
    from full_waveform_inversion_b200 import io as fio
    hyp = tmp_path / "loc.hyp"
    hyp.write_text(HYP)
    uid, stations = fio.get_event_uid_and_station_data_MTFIT_FORMAT_from_nonlinloc_hyp_file(str(hyp))
    assert uid == "20180214185538216400"
    assert [s[0][0] for s in stations] == ["RA51", "RA52"]
    assert stations[0][1][0, 0] == 123.4 and abs(stations[0][2][0, 0] - (180.0 - 35.0)) < 1e-12 and stations[0][3][0, 0] == 0
    MTs, MTp = np.arange(21.0).reshape(7, 3), np.array([0.2, 0.0, 0.8])
    f = fio.save_to_MTFIT_style_file(MTs, MTp, str(hyp), "DC_crack_couple", str(tmp_path), MTp_absolute=np.ones(3))
    assert f.endswith("20180214185538216400_FW_DC_crack_couple.pkl")
    d = pickle.load(open(f, "rb"))
    assert set(d) == {"MTs", "MTp", "uid", "stations", "MTp_absolute"}                       # FWI:961-967
    uid2, MTp2, MTs2, st2 = fio.load_MT_dict_from_file(f)                                    # the reader PLOT / UNP use
    assert uid2 == uid and np.array_equal(MTs2, MTs)
    w = fio.save_specific_waveforms_to_file(np.ones((2, 5)), np.zeros((2, 5)), ["RA51, L", "RA52, L"], str(hyp), "DC", str(tmp_path))
    wd = pickle.load(open(w, "rb"))
    assert set(wd) == {"RA51, L", "RA52, L"} and set(wd["RA51, L"]) == {"real_wf", "synth_wf"}   # FWI:1026-1029
    p2, s2 = fio.remove_zero_prob_results(MTp, MTs)
    assert np.array_equal(p2, [0.2, 0.8]) and s2.shape == (7, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_loaders_bit_exact_vs_reference(golden_a, tmp_path, case):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from full_waveform_inversion_b200 import io as fio
    g = golden_a
    names = _write_files(str(tmp_path), g)
    itype, kwf = CASES[case]
    kw = kwf(g)
    multi = kw.get("invert_for_ratio_of_multiple_media_greens_func_switch", False)
    real, G = fio.get_overall_real_and_green_func_data(str(tmp_path), names["real"], names["mt"] + (names["mt2"] if multi else []),
                                                       names["sf"] + (names["sf2"] if multi else []), itype, **kw)
    assert G.dtype == np.float64 and G.shape == g[case + "_G"].shape
    assert np.array_equal(G, g[case + "_G"])                     # roll / zero / cut / scale are exact operations
    assert np.array_equal(real, g[case + "_real"])


@pytest.mark.gpu
def test_run_end_to_end(golden_a, tmp_path):
    """The drop-in `run`: files in -> LSQ + Monte-Carlo -> the reference's .pkl / .wfs files out -> UNP re-evaluation."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from full_waveform_inversion_b200 import io as fio
    K, T = 6, 128
    d, G, m_true = orc.synthetic_inputs(K=K, C=9, T=T, seed=4)
    G = G * 1e-7                                                 # so that the loaders' unit factors (x 1e7) bring it back to O(1)
    data_dir = tmp_path / "data"
    data_dir.mkdir()
    names = {"real": [], "mt": [], "sf": []}
    for k in range(K):
        np.savetxt(data_dir / ("real_%d.txt" % k), d[k], fmt="%.17e")
        np.savetxt(data_dir / ("mt_%d.txt" % k), (G[k, :6] / 1e3).T, fmt="%.17e")
        np.savetxt(data_dir / ("sf_%d.txt" % k), G[k, 6:].T, fmt="%.17e")
        for key in names:
            names[key].append("%s_%d.txt" % (key, k))
    hyp = tmp_path / "loc.hyp"
    hyp.write_text(HYP)
    out = tmp_path / "out"
    labels = ["S%d, Z" % k for k in range(K)]
    MTs, MTp, MTp_abs = fio.run(str(data_dir), str(out), names["real"], names["mt"], names["sf"], labels,
                                "single_force_crack_no_coupling", False, False, 4000, "VR", [], [], str(hyp),
                                return_absolute_similarity_values_switch=True, seed=3)
    assert MTs.shape == (10, 4000) and abs(MTp.sum() - 1.0) < 1e-5
    uid = "20180214185538216400"
    for sub in ("", "least_squares_result/"):
        assert os.path.exists(out / (sub + uid + "_FW_single_force_crack_no_coupling.pkl"))
        assert os.path.exists(out / (sub + uid + "_FW_single_force_crack_no_coupling.wfs"))
    uid2, MTp2, MTs2, _ = fio.load_MT_dict_from_file(str(out / (uid + "_FW_single_force_crack_no_coupling.pkl")))
    assert np.array_equal(MTs2, MTs) and np.array_equal(MTp2, MTp)
    # UNP: similarity of the most likely sample == inverse of the likelihood transform of its absolute probability
    s = fio.unnormallised_probability_run("single_force_crack_no_coupling", uid, str(out), str(data_dir), names["real"],
                                          names["mt"], names["sf"], [], "VR", False, False)
    L_best = MTp_abs[int(np.argmax(MTp))]
    assert abs(s - (1.0 + 2.0 * np.log(L_best))) < 2e-5
    wfs = pickle.load(open(out / (uid + "_FW_single_force_crack_no_coupling.wfs"), "rb"))
    assert set(wfs) == set(labels)
    best = MTs[:9, int(np.argmax(MTp))]
    want = orc.forward_model(G * 1e7, best)                      # the loaders' unit factors: files hold G (MT / 1e3), loaded = G * 1e7
    got = np.stack([wfs[lab]["synth_wf"] for lab in labels])
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("with_phase_labels", [False, True])
def test_run_multi_medium_inversion_end_to_end(tmp_path, with_phase_labels):
    """`run_multi_medium_inversion` (FWI:1037-1158): two sets of Green's-function files -> (K,C,T,2) -> LSQ on the 50/50 mix
    (FWI:1059-1063) -> Monte-Carlo with media-ratio rows -> files; the most likely waveforms use the sample's own ratio(s)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from full_waveform_inversion_b200 import io as fio
    K, T = 6, 96
    d, G2, _ = orc.synthetic_inputs(K=K, C=9, T=T, seed=9, n_media=2)
    G2 = G2 * 1e-7
    data_dir = tmp_path / "data"
    data_dir.mkdir()
    names = {"real": [], "mt": [], "sf": []}
    for k in range(K):
        np.savetxt(data_dir / ("real_%d.txt" % k), d[k], fmt="%.17e")
        names["real"].append("real_%d.txt" % k)
    for medium in (0, 1):                                        # medium-1 file names first, then medium 2 (split index = K)
        for k in range(K):
            np.savetxt(data_dir / ("mt_%d_m%d.txt" % (k, medium)), (G2[k, :6, :, medium] / 1e3).T, fmt="%.17e")
            np.savetxt(data_dir / ("sf_%d_m%d.txt" % (k, medium)), G2[k, 6:, :, medium].T, fmt="%.17e")
            names["mt"].append("mt_%d_m%d.txt" % (k, medium))
            names["sf"].append("sf_%d_m%d.txt" % (k, medium))
    hyp = tmp_path / "loc.hyp"
    hyp.write_text(HYP)
    out = tmp_path / "out"
    labels = ["S%d, Z" % k for k in range(K)]
    phases = ["P", "S", "surface", "P", "S", "S"] if with_phase_labels else []
    itype = "single_force_crack_no_coupling"
    MTs, MTp, MTp_abs = fio.run_multi_medium_inversion(
        str(data_dir), str(out), names["real"], names["mt"], names["sf"], labels, itype, False, False, 3000, "VR", [], [], str(hyp),
        return_absolute_similarity_values_switch=True, green_func_fnames_split_index=K, green_func_phase_labels=phases, seed=5)
    nfrac = 3 if with_phase_labels else 1
    assert MTs.shape == (9 + 1 + nfrac, 3000) and abs(MTp.sum() - 1.0) < 1e-5
    uid = "20180214185538216400"
    assert os.path.exists(out / (uid + "_FW_%s.pkl" % itype)) and os.path.exists(out / ("least_squares_result/" + uid + "_FW_%s.pkl" % itype))
    j = int(np.argmax(MTp))
    fr = MTs[10:, j]
    G_loaded = G2 * 1e7
    pidx = np.array([("P", "S", "surface").index(x) for x in phases]) if with_phase_labels else None
    want = orc.forward_model(orc.mix_media(G_loaded, fr if with_phase_labels else fr[0], pidx), MTs[:9, j])
    wfs = pickle.load(open(out / (uid + "_FW_%s.wfs" % itype), "rb"))
    got = np.stack([wfs[lab]["synth_wf"] for lab in labels])
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-5
    # the sample's likelihood is what the oracle gives for that source and those ratios
    s = orc.compare_synth_to_real_waveforms(d, want, "VR", False, False)
    assert abs(MTp_abs[j] - orc.likelihood(s)) <= 2e-5
