"""ctypes binding of include/wf_engine.h.  Fails loudly if the CUDA library is missing — there is
no CPU fallback and nothing under oracle/ is ever imported from here."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libwf_b200.so")


class wf_material(C.Structure):
    _fields_ = [("model", C.c_int), ("E", C.c_double), ("nu", C.c_double), ("rho0", C.c_double),
                ("sy0", C.c_double), ("K", C.c_double), ("m", C.c_double), ("q", C.c_double * 14),
                ("temp", C.c_double), ("max_edot", C.c_double)]


STAB_FIELDS = ("alpha_free alpha_contact hg_coeff_free hg_coeff_contact av_coeff_div av_coeff_bulk "
               "log_factor pspg_scale p_pspg_bulkfac J_min hg_visc hg_stiff hexa_hg_coeff").split()


class wf_stab(C.Structure):
    _fields_ = [(f, C.c_double) for f in STAB_FIELDS]


_lib = None

UNFUSED = ("UpdatePrediction calcElemJAndDerivatives Calc_Element_Radius CalcElemVol CalcNodalVol "
           "CalcNodalMassFromVol calcElemStrainRates calcElemPressure calcArtificialViscosity calcElemForces "
           "calcElemHourglassForces assemblyForces calcAccel UpdateCorrectionAccVel AxisConstraint "
           "UpdateCorrectionPos SearchExtNodes CalcExtFaceAreas CalcContactForces MoveTriMesh").split()

# every symbol include/wf_engine.h declares (checked by tests/test_abi.py)
DECLARED = (["wf_create", "wf_destroy", "wf_last_error", "wf_set_stream", "wf_get_stream", "wf_synchronize", "wf_set_mesh", "wf_gen_box",
             "wf_get_counts", "wf_set_axisymm_vol_weight", "wf_set_material", "wf_set_stab", "wf_set_options",
             "wf_set_tracking", "wf_add_bc_vel", "wf_add_bc_vel_array", "wf_allocate_bcs", "wf_set_bc_values", "wf_init", "wf_step",
             "wf_nonfinite_flag", "wf_energies", "wf_monitor_async", "wf_monitor_wait", "wf_calcMinEdgeLength", "wf_max_velocity", "wf_cfl_dt", "wf_set_dt", "wf_get_time", "wf_set_time", "wf_step_timed", "wf_set_variant", "wf_ImposeBCV", "wf_ImposeBCA", "wf_CalcStressStrain",
             "wf_get_array", "wf_set_array", "wf_array_bytes", "wf_device_ptr", "wf_partition_build",
             "wf_partition_build_box", "wf_partition_free", "wf_partition_info", "wf_partition_node_l2g",
             "wf_partition_local_elnod", "wf_partition_neigh_ranks", "wf_partition_halo_offset",
             "wf_partition_halo_nodes", "wf_set_mesh_partition", "wf_set_axis_xmin", "wf_step_phase", "wf_init_phase", "wf_halo_info",
             "wf_halo_comm_block", "wf_halo_slot_offsets", "wf_halo_ipc_export", "wf_halo_ipc_open", "wf_halo_connect",
             "wf_halo_status", "wf_connect_all", "wf_init_all", "wf_step_all", "wf_halo_set_transport", "wf_halo_exchange_ptrs",
             "wf_host_box_counts", "wf_host_gen_box", "wf_host_nodel", "wf_version",
             "wf_set_trimesh", "wf_set_contact", "wf_get_trimesh_counts", "wf_host_ext_faces",
             "wf_host_axis_plane_counts", "wf_host_axis_plane_mesh", "wf_set_thermal", "wf_set_contact_heat",
             "wf_host_force_tiles", "wf_set_elem_order", "wf_host_elem_order", "wf_host_elem_order_keys", "wf_host_brick_plan", "wf_brick_info", "wf_host_run_slots", "wf_step_open", "wf_step_close"]
            + ["wf_" + n for n in UNFUSED])


def load():
    """Load libwf_b200.so (building it first if the sources are newer and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m weldformfem_b200.build` "
            "(nvcc, sm_100a). The engine has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, ip, dp, up = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_uint)
    sig = {
        "wf_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int]),
        "wf_destroy": (None, [vp]),
        "wf_last_error": (C.c_char_p, [vp]),
        "wf_set_stream": (C.c_int, [vp, vp]),
        "wf_get_stream": (C.c_int, [vp, C.POINTER(vp)]),
        "wf_synchronize": (C.c_int, [vp]),
        "wf_set_mesh": (C.c_int, [vp, C.c_int, C.c_int, dp, up]),
        "wf_gen_box": (C.c_int, [vp, dp, dp, C.c_double, C.c_int]),
        "wf_get_counts": (C.c_int, [vp, ip, ip, ip]),
        "wf_set_axisymm_vol_weight": (C.c_int, [vp, C.c_int]),
        "wf_set_material": (C.c_int, [vp, C.POINTER(wf_material)]),
        "wf_set_stab": (C.c_int, [vp, C.POINTER(wf_stab)]),
        "wf_set_options": (C.c_int, [vp, C.c_int, C.c_double, C.c_double, C.c_int]),
        "wf_set_tracking": (C.c_int, [vp, C.c_int]),
        "wf_add_bc_vel": (C.c_int, [vp, C.c_int, C.c_int, C.c_double]),
        "wf_add_bc_vel_array": (C.c_int, [vp, C.c_int, ip, ip, dp]),
        "wf_allocate_bcs": (C.c_int, [vp]),
        "wf_set_bc_values": (C.c_int, [vp, C.c_int, C.c_int, dp]),
        "wf_init": (C.c_int, [vp, C.c_double]),
        "wf_init_phase": (C.c_int, [vp, C.c_int, C.c_double]),
        "wf_step": (C.c_int, [vp, C.c_int]),
        "wf_step_open": (C.c_int, [vp, C.c_int]),
        "wf_step_close": (C.c_int, [vp]),
        "wf_step_phase": (C.c_int, [vp, C.c_int, C.c_int]),
        "wf_step_timed": (C.c_int, [vp, C.c_int, C.POINTER(C.c_float)]),
        "wf_set_variant": (C.c_int, [vp, C.c_int, C.c_int]),
        "wf_nonfinite_flag": (C.c_int, [vp, ip]),
        "wf_energies": (C.c_int, [vp, dp, dp]),
        "wf_calcMinEdgeLength": (C.c_int, [vp, dp, dp]),
        "wf_max_velocity": (C.c_int, [vp, dp]),
        "wf_cfl_dt": (C.c_int, [vp, C.c_double, dp]),
        "wf_set_dt": (C.c_int, [vp, C.c_double]),
        "wf_monitor_async": (C.c_int, [vp]),
        "wf_monitor_wait": (C.c_int, [vp, dp, ip]),
        "wf_get_time": (C.c_int, [vp, dp, C.POINTER(C.c_long)]),
        "wf_set_time": (C.c_int, [vp, C.c_double, C.c_long]),
        "wf_ImposeBCV": (C.c_int, [vp, C.c_int]),
        "wf_ImposeBCA": (C.c_int, [vp, C.c_int]),
        "wf_CalcStressStrain": (C.c_int, [vp, C.c_double]),
        "wf_get_array": (C.c_int, [vp, C.c_char_p, vp, C.c_size_t]),
        "wf_set_array": (C.c_int, [vp, C.c_char_p, vp, C.c_size_t]),
        "wf_array_bytes": (C.c_size_t, [vp, C.c_char_p]),
        "wf_device_ptr": (vp, [vp, C.c_char_p, C.POINTER(C.c_size_t)]),
        "wf_halo_info": (C.c_int, [vp, ip, ip, ip, C.POINTER(ip), C.POINTER(ip), C.POINTER(ip)]),
        "wf_halo_comm_block": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "wf_halo_slot_offsets": (C.c_int, [vp, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
        "wf_halo_ipc_export": (C.c_int, [vp, vp]),
        "wf_halo_ipc_open": (C.c_int, [vp, vp, C.POINTER(vp)]),
        "wf_halo_connect": (C.c_int, [vp, C.c_int, vp, C.c_size_t, C.c_size_t]),
        "wf_halo_status": (C.c_int, [vp, ip]),
        "wf_connect_all": (C.c_int, [C.POINTER(vp), C.c_int]),
        "wf_init_all": (C.c_int, [C.POINTER(vp), C.c_int, C.c_double]),
        "wf_step_all": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int]),
        "wf_halo_set_transport": (C.c_int, [vp, C.c_int]),
        "wf_halo_exchange_ptrs": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "wf_partition_build": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, up]),
        "wf_partition_build_box": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, dp, dp, C.c_double, C.c_int]),
        "wf_partition_free": (None, [vp]),
        "wf_partition_info": (C.c_int, [vp, ip, ip, ip, ip]),
        "wf_partition_node_l2g": (ip, [vp]),
        "wf_partition_local_elnod": (up, [vp]),
        "wf_partition_neigh_ranks": (ip, [vp]),
        "wf_partition_halo_offset": (ip, [vp]),
        "wf_partition_halo_nodes": (ip, [vp]),
        "wf_set_mesh_partition": (C.c_int, [vp, vp, dp]),
        "wf_set_axis_xmin": (C.c_int, [vp, C.c_double]),
        "wf_host_box_counts": (C.c_int, [dp, C.c_double, C.c_int, ip, ip, ip, ip]),
        "wf_host_gen_box": (C.c_int, [dp, dp, C.c_double, C.c_int, dp, up]),
        "wf_host_nodel": (C.c_int, [C.c_int, C.c_int, C.c_int, up, ip, ip, ip, ip]),
        "wf_version": (C.c_char_p, []),
        "wf_set_trimesh": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, dp, dp, ip, dp, ip]),
        "wf_set_contact": (C.c_int, [vp, C.c_double, C.c_double, C.c_double, C.c_double]),
        "wf_get_trimesh_counts": (C.c_int, [vp, ip, ip, ip]),
        "wf_set_thermal": (C.c_int, [vp] + [C.c_double] * 5),
        "wf_set_contact_heat": (C.c_int, [vp, C.c_double, C.c_double]),
        "wf_host_ext_faces": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, up, C.POINTER(C.c_ubyte), ip, ip, ip, ip]),
        "wf_host_axis_plane_counts": (C.c_int, [C.c_int, C.c_int, ip, ip]),
        "wf_host_force_tiles": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, up, C.POINTER(C.c_longlong), C.POINTER(C.c_ubyte),
                                          C.POINTER(C.c_longlong), up, C.POINTER(C.c_ubyte)]),
        "wf_set_elem_order": (C.c_int, [vp, C.c_int]),
        "wf_host_run_slots": (C.c_int, [C.c_int, ip, ip]),
        "wf_host_elem_order": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, dp, up, C.c_int, ip]),
        "wf_host_elem_order_keys": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, dp, up, C.c_int, ip, C.POINTER(C.c_ulonglong)]),
        "wf_brick_info": (C.c_int, [vp, ip, ip]),
        "wf_host_brick_plan": (C.c_int, [C.c_int, C.POINTER(C.c_ulonglong), ip, ip]),
        "wf_host_axis_plane_mesh": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, dp, ip, dp, ip]),
    }
    for n in UNFUSED:
        sig["wf_" + n] = (C.c_int, [vp])
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    _lib = lib
    return lib
