// pv_march.cuh -- records exchanged between the march kernels (pv_march.cu) and the gather kernel (pv_gather.cu).
//
// PhotonVolumeIntegrator::Li (integrators/photonvolume.cpp:112-222) is split in two on the device:
//   march  -- everything of a march step that does not depend on the Lv/Tr recurrence or on the photon map:
//             sample position, optical depth of the step segment, Russian-roulette draw, density at the sample,
//             light choice, shadow ray (LinearBVHNode walk) and its optical depth.  One thread per step.
//   gather -- the photon lookup + radiance estimate of each step and the recurrence, one warp per ray.
// Seven scalars per step cross from one to the other in a 32-byte StepRec; a 32-byte RayHdr per ray says where
// the ray's steps are.
#pragma once
#include <stdint.h>

struct RayHdr {
    unsigned long long offset;   // first StepRec of the ray
    int nSamples;                // 0: the ray misses the medium (L = 0, T = 1)
    float step;                  // (t1 - t0) / nSamples
    float pad[4];
};
struct StepRec {
    float t;      // sample parameter: p = ray(t), t accumulated as the reference does (t0 += step)
    float tau;    // optical-depth scalar of the step segment: tau[b] = sigma_t[b] * tau
    float rr;     // Russian-roulette draw if Tr.y() < 1e-3 at this step, else -1; PV_RR_DEAD: the march ended at an earlier step
    float dens;   // density at p (1/0 inside/outside for homogeneous media)
    float sh;     // optical-depth scalar of the shadow ray
    float dfac;   // falloff / dist^2 * phase * nLights (0: unlit or occluded): L_d[b] = I[b] * exp(-sigma_t[b]*sh) * dfac
    int ln;       // light chosen for the step
    uint32_t pad; // index of the step's ray within the slice (the step-parallel gather finds its ray through it)
};
// rr of the records behind a step whose roulette draw ended the march (draw > continueProb = .5, photonvolume.cpp:160-163):
// the recurrence never reaches them, the step-sorted lookups skip them
#define PV_RR_DEAD 3.f
static_assert(sizeof(RayHdr) == 32 && sizeof(StepRec) == 32, "march records are two 128-bit words");
