"""ctypes mirror of include/pv.h (the C ABI of the photon-volume path).

Field order and types must match include/pv.h exactly; tests/test_abi.py checks
the struct sizes against the values the C compiler reports.
"""
import ctypes as C

NSPEC = 30
PV_OK, PV_EINVAL, PV_ECUDA, PV_ENOMEM, PV_ESTATE, PV_ENOPHOTONS = 0, -1, -2, -3, -4, -5
LIGHT_POINT, LIGHT_SPOT, LIGHT_DISTANT, LIGHT_AREA = 0, 1, 2, 3
MEDIUM_NONE, MEDIUM_HOMOGENEOUS, MEDIUM_GRID, MEDIUM_RAINBOW, MEDIUM_EXPONENTIAL = 0, 1, 2, 3, 4
MAT_MATTE, MAT_GLASS = 0, 1
GATHER_NO_DIRECT, GATHER_NO_INDIRECT, GATHER_RAY_PARALLEL, GATHER_STEP_PARALLEL, GATHER_CELL_BATCHED = 1, 2, 4, 8, 16

Spec = C.c_float * NSPEC
Mat16 = C.c_float * 16


class BvhNode(C.Structure):
    _fields_ = [("bounds", C.c_float * 6), ("offset", C.c_uint32), ("n_primitives", C.c_uint8),
                ("axis", C.c_uint8), ("pad", C.c_uint8 * 2)]


class Ray(C.Structure):
    _fields_ = [("o", C.c_float * 3), ("d", C.c_float * 3), ("mint", C.c_float), ("maxt", C.c_float),
                ("time", C.c_float), ("u_scatter", C.c_float)]


class Light(C.Structure):
    _fields_ = [("type", C.c_int32), ("pos", C.c_float * 3), ("dir", C.c_float * 3),
                ("cos_total_width", C.c_float), ("cos_falloff_start", C.c_float), ("intensity", Spec),
                ("light_to_world", Mat16), ("world_to_light", Mat16), ("power_y", C.c_float)]


class Medium(C.Structure):
    _fields_ = [("type", C.c_int32), ("world_to_volume", Mat16), ("p0", C.c_float * 3), ("p1", C.c_float * 3),
                ("sigma_a", Spec), ("sigma_s", Spec), ("le", Spec), ("g", C.c_float),
                ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("density", C.POINTER(C.c_float))]


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("kd", Spec), ("kr", Spec), ("kt", Spec), ("index", C.c_float), ("vn", C.c_float)]


class Sphere(C.Structure):
    _fields_ = [("object_to_world", Mat16), ("world_to_object", Mat16), ("radius", C.c_float), ("zmin", C.c_float), ("zmax", C.c_float),
                ("theta_min", C.c_float), ("theta_max", C.c_float), ("phi_max", C.c_float), ("flip_normal", C.c_int32)]


SHAPE_TRIANGLE = 0xFFFFFFFF


class SceneDesc(C.Structure):
    _fields_ = [("nodes", C.POINTER(BvhNode)), ("n_nodes", C.c_uint32),
                ("tri_verts", C.POINTER(C.c_float)), ("prim_material", C.POINTER(C.c_uint32)), ("n_prims", C.c_uint32),
                ("materials", C.POINTER(Material)), ("n_materials", C.c_uint32),
                ("lights", C.POINTER(Light)), ("n_lights", C.c_uint32),
                ("medium", C.POINTER(Medium)), ("world_bound", C.c_float * 6), ("cie_y", Spec),
                ("prim_shape", C.POINTER(C.c_uint32)), ("spheres", C.POINTER(Sphere)), ("n_spheres", C.c_uint32),
                ("light_tris", C.POINTER(C.c_float)), ("n_light_tris", C.c_uint32)]


class GatherParams(C.Structure):
    _fields_ = [("stepsize", C.c_float), ("nused", C.c_uint32), ("maxdist", C.c_float), ("seed", C.c_uint64),
                ("ray_index_base", C.c_uint64), ("flags", C.c_uint32)]


class ShootParams(C.Structure):
    _fields_ = [("stepsize", C.c_float), ("integrator_stepsize", C.c_float), ("max_photon_depth", C.c_int32),
                ("seed", C.c_uint64), ("rank", C.c_uint32), ("world", C.c_uint32), ("max_paths", C.c_uint64),
                ("time", C.c_float)]


class ShootStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("paths_local", C.c_uint64), ("photons_local", C.c_uint64),
                ("blocks", C.c_uint64), ("nodes_visited", C.c_uint64), ("tri_tests", C.c_uint64),
                ("density_samples", C.c_uint64), ("segments", C.c_uint64), ("stack_overflows", C.c_uint64),
                ("seconds", C.c_double)]


class GatherStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("lookups", C.c_uint64), ("photons_found", C.c_uint64),
                ("candidates_tested", C.c_uint64), ("heap_lookups", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("density_samples", C.c_uint64)]


MAP_VOLUME, MAP_CAUSTIC, MAP_INDIRECT, MAP_DIRECT, MAP_RADIANCE = 0, 1, 2, 3, 4


class MapsParams(C.Structure):
    _fields_ = [("n_volume_wanted", C.c_uint64), ("n_caustic_wanted", C.c_uint64), ("n_indirect_wanted", C.c_uint64),
                ("final_gather", C.c_int32)]


class MapsStats(C.Structure):
    _fields_ = [("nshot", C.c_uint64), ("blocks", C.c_uint64), ("n_caustic_paths", C.c_uint64), ("n_indirect_paths", C.c_uint64),
                ("n_direct_paths", C.c_uint64), ("n_volume_paths", C.c_uint64), ("n", C.c_uint64 * 5),
                ("replayed_blocks", C.c_uint64), ("shoot", ShootStats)]


ALLREDUCE_U32_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint32), C.c_uint64, C.c_void_p)


VOLINT_SINGLE, VOLINT_EMISSION = 0, 1        # PV_VOLINT_*
VOLINT_WARP_PER_RAY, VOLINT_THREAD_PER_RAY = 4, 8

# every symbol include/pv.h declares (tests/test_abi.py checks the .so exports them all)
EXPORTS = [
    "pv_create", "pv_destroy", "pv_last_error", "pv_version", "pv_set_scene", "pv_set_photons",
    "pv_set_photons_dev", "pv_get_photons", "pv_get_photons_dev", "pv_photon_count", "pv_build", "pv_knn",
    "pv_intersect", "pv_occluded", "pv_transmittance", "pv_gather", "pv_gather_dev", "pv_lphoton",
    "pv_gather_stats_get", "pv_last_kernel_ms", "pv_last_march_ms", "pv_last_phase_ms", "pv_launch_count", "pv_comm_unique_id", "pv_comm_init", "pv_comm_init_all", "pv_comm_destroy", "pv_allgather_photons", "pv_broadcast_photons", "pv_shoot", "pv_shoot_blocks", "pv_shoot_finish", "pv_stream",
    "pv_shoot_maps", "pv_shoot_maps_ranks", "pv_get_map_photons", "pv_set_map_photons", "pv_radiance_photons", "pv_select_map", "pv_surface_lphoton", "pv_radiance_nearest", "pv_final_gather", "pv_set_radiance_lo",
    "pv_volume_li", "pv_volume_li_dev", "pv_gather_indexed", "pv_volume_li_indexed", "pv_build_bvh",
]
