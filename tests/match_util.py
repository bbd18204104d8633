"""Shared helpers of the matchToMap tests: oracle plumbing on the flat map contract of alvaar_b200.synth.make_match_problem."""
import ctypes as C

import numpy as np

from conftest import P

_vp, _i, _d, _f = C.c_void_p, C.c_int, C.c_double, C.c_float
ORC_ARGS = [_i, _i, _d, _d, _d, _d, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _f, _f, _vp, _vp]
REF_ARGS = ORC_ARGS[:28] + [_vp, _vp, _vp]


def oracle_match(oracle, p, order, nkp3d, max_err=2.0, ratio=0.2):
    oracle.orc_match_to_map.argtypes = ORC_ARGS
    mk, mm = np.zeros(len(p["kp_id"]) + 1, np.int32), np.zeros(len(p["kp_id"]) + 1, np.int32)
    order = np.ascontiguousarray(order, np.int32)
    n = oracle.orc_match_to_map(p["w"], p["h"], p["K"][0], p["K"][1], p["K"][2], p["K"][3], P(p["cur_T"]), len(p["kp_id"]), P(p["kp_id"]),
                                P(p["kp_px"]), nkp3d, len(p["kf_id"]), P(p["kf_id"]), P(p["kf_T"]), len(p["mp_id"]), P(p["mp_id"]),
                                P(p["mp_wpt"]), P(p["mp_is3d"]), P(p["obs_start"]), P(p["obs_kf"]), P(p["obs_px"]), P(p["desc_start"]),
                                P(p["desc_kf"]), P(p["desc"]), len(order), P(order), max_err, ratio, P(mk), P(mm))
    return dict(zip(mk[:n].tolist(), mm[:n].tolist()))


def reference_match(ref, p, nkp3d, max_err=2.0, ratio=0.2):
    ref.ref_match_to_map.argtypes = REF_ARGS
    nl = len(p["local_ids"])
    order, mk, mm = np.zeros(nl, np.int32), np.zeros(len(p["kp_id"]) + 1, np.int32), np.zeros(len(p["kp_id"]) + 1, np.int32)
    n = ref.ref_match_to_map(p["w"], p["h"], p["K"][0], p["K"][1], p["K"][2], p["K"][3], P(p["cur_T"]), len(p["kp_id"]), P(p["kp_id"]),
                             P(p["kp_px"]), nkp3d, len(p["kf_id"]), P(p["kf_id"]), P(p["kf_T"]), len(p["mp_id"]), P(p["mp_id"]),
                             P(p["mp_wpt"]), P(p["mp_is3d"]), P(p["obs_start"]), P(p["obs_kf"]), P(p["obs_px"]), P(p["desc_start"]),
                             P(p["desc_kf"]), P(p["desc"]), nl, P(p["local_ids"]), max_err, ratio, P(order), P(mk), P(mm))
    return order, dict(zip(mk[:n].tolist(), mm[:n].tolist()))
