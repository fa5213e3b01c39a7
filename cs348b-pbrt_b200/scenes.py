"""Synthetic workloads named by BASELINE.json configs 2, 3 and 5.

The geometry/light/medium SPECTRA of these scenes come from the reference's own
RGB->spectrum conversion, so they are exported once by oracle/ref_harness
(--export-scene) into tests/golden/*.scn; this module only adds what is closed
form: the density grid, the camera rays and the synthetic photon sets.
"""
import numpy as np

CORNELL_TEMPLATE = """# {title}
Film "image" "string filename" "{outfile}"
 "integer xresolution" [{xres}] "integer yresolution" [{yres}]
Sampler "lowdiscrepancy" "integer pixelsamples" [1]
PixelFilter "box"
SurfaceIntegrator "photonmap" "integer nused" [50] "bool finalgather" ["false"]
  "float maxdist" [.1] "integer indirectphotons" [0] "integer causticphotons" [0]
  "float stepsize" [{shoot_step}] "integer maxphotondepth" [5]
VolumeIntegrator "photonvolume" "float stepsize" [{stepsize}] "integer nused" [{nused}] "float maxdist" [{maxdist}]
  "integer volumephotons" [{nphotons}]
LookAt 0 0 -3.4  0 0 0  0 1 0
Camera "perspective" "float fov" [40]
WorldBegin
{volume}
LightSource "point" "point from" [0 0.8 0] "color I" [20 20 20]
Material "matte" "color Kd" [.6 .6 .6]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-1 -1 -1  1 -1 -1  1 -1 1  -1 -1 1]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-1 1 -1  1 1 -1  1 1 1  -1 1 1]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-1 -1 1  1 -1 1  1 1 1  -1 1 1]
Material "matte" "color Kd" [.6 .1 .1]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-1 -1 -1  -1 -1 1  -1 1 1  -1 1 -1]
Material "matte" "color Kd" [.1 .6 .1]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [1 -1 -1  1 -1 1  1 1 1  1 1 -1]
Material "matte" "color Kd" [.6 .6 .6]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-1 -1 -4  1 -1 -4  1 1 -4  -1 1 -4]
WorldEnd
"""

# The same box with every photon map of the reference's shooter switched on: a dispersive glass wedge under
# the light makes real caustics, the matte walls indirect / direct / radiance photons, the medium volume photons.
CORNELL_SURF_TEMPLATE = """# synthetic Cornell box, all photon maps on (surface maps: core/photonshooter.cpp:147-189)
Film "image" "string filename" "{outfile}"
 "integer xresolution" [{xres}] "integer yresolution" [{yres}]
Sampler "lowdiscrepancy" "integer pixelsamples" [1]
PixelFilter "box"
SurfaceIntegrator "photonmap" "integer nused" [{surf_nused}] "bool finalgather" ["{finalgather}"] "integer finalgathersamples" [{fgsamples}]
  "float maxdist" [{surf_maxdist}] "integer indirectphotons" [{indirect}] "integer causticphotons" [{caustic}]
  "float stepsize" [{shoot_step}] "integer maxphotondepth" [5]
VolumeIntegrator "photonvolume" "float stepsize" [{stepsize}] "integer nused" [{nused}] "float maxdist" [{maxdist}]
  "integer volumephotons" [{nphotons}]
LookAt 0 0 -3.4  0 0 0  0 1 0
Camera "perspective" "float fov" [40]
WorldBegin
{volume}
LightSource "point" "point from" [0 0.8 0] "color I" [20 20 20]
Material "matte" "color Kd" [.6 .6 .6]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-1 -1 -1  1 -1 -1  1 -1 1  -1 -1 1]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-1 1 -1  1 1 -1  1 1 1  -1 1 1]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-1 -1 1  1 -1 1  1 1 1  -1 1 1]
Material "matte" "color Kd" [.6 .1 .1]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-1 -1 -1  -1 -1 1  -1 1 1  -1 1 -1]
Material "matte" "color Kd" [.1 .6 .1]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [1 -1 -1  1 -1 1  1 1 1  1 1 -1]
Material "matte" "color Kd" [.6 .6 .6]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-1 -1 -4  1 -1 -4  1 1 -4  -1 1 -4]
AttributeBegin
Material "glass" "float index" [1.5] "float Vn" [{vn}] "color Kr" [1 1 1] "color Kt" [1 1 1]
Translate 0 -0.3 0.1
Scale 0.45 0.25 0.45
Shape "trianglemesh" "point P" [1 -1 -1  1 -1 1  -1 -1 1  -1 -1 -1  1 1 0  -1 1 0]
  "integer indices" [0 1 2  0 2 3  1 4 5  1 5 2  0 4 1  2 5 3  4 0 3  4 3 5]
AttributeEnd
WorldEnd
"""

HOMOG_VOLUME = ('Volume "homogeneous" "color sigma_a" [.3 .3 .3] "color sigma_s" [.15 .15 .15] "float g" [0]\n'
                '  "point p0" [-1 -1 -1] "point p1" [1 1 1]')


# ExponentialDensity (volumes/exponential.h): density a * exp(-b * height above the extent's floor)
EXP_VOLUME = ('Volume "exponential" "color sigma_a" [.3 .3 .3] "color sigma_s" [.6 .6 .6] "float g" [0.2]\n'
              '  "point p0" [-1 -1 -1] "point p1" [1 1 1] "float a" [2] "float b" [1.5] "vector updir" [0.2 1 0.1]')


# BASELINE config 1: the parameters of the reference's projectScene/volumescene_png.pbrt (rainbow medium, distant light,
# three matte quads, photonmap surface integrator with final gathering), with the counts / resolution / output as knobs.
VOLUMESCENE_TEMPLATE = """# BASELINE.json configs[0]: rainbow-volume scene of the reference project
Film "image" "string filename" "{outfile}" "integer xresolution" [{xres}] "integer yresolution" [{yres}]
Sampler "lowdiscrepancy" "integer pixelsamples" [{spp}]
PixelFilter "{filt}"
SurfaceIntegrator "photonmap" "integer nused" [300] "bool finalgather" ["{finalgather}"] "integer finalgathersamples" [64]
  "float maxdist" [.15] "integer indirectphotons" [0] "integer causticphotons" [{caustic}]
VolumeIntegrator "photonvolume" "float stepsize" [.15] "integer nused" [50] "float maxdist" [0.5]
  "integer volumephotons" [{nphotons}]
Rotate 0 1 0 0
Camera "perspective" "float fov" [70]
WorldBegin
Translate 0 -0.5 3.5
Volume "rainbow" "color sigma_a" [.05 .05 .05] "color sigma_s" [.1 .1 .1] "point p0" [-10 0 -5] "point p1" [5 5 5]
AttributeBegin
LightSource "distant" "point from" [0 3 0] "point to" [0 2 5] "color L" [150 150 150]
AttributeEnd
Material "matte" "color Kd" [.01 .01 .01]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-5 0 -5  5 0 -5  5 0 5  -5 0 5]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-5 0 3  5 0 3  5 10 3  -5 10 3]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [5 0 3  5 0 -3  5 10 -3  5 10 3]
WorldEnd
"""

# BASELINE config 4: the parameters of projectScene/pinkfloyd.pbrt (homogeneous medium, narrow spot + point light, a
# dispersive glass wedge -- 8 triangles -- and a matte wall), wedge mesh written out inline.
PRISM_TEMPLATE = """# BASELINE.json configs[3]: prism-dispersion scene of the reference project
Film "image" "integer xresolution" [{xres}] "integer yresolution" [{yres}] "string filename" "{outfile}"
Sampler "lowdiscrepancy" "integer pixelsamples" [{spp}]
PixelFilter "{filt}"
SurfaceIntegrator "photonmap" "integer nused" [50] "bool finalgather" ["false"] "float maxdist" [.15]
  "integer indirectphotons" [0] "integer causticphotons" [{caustic}]
VolumeIntegrator "photonvolume" "float stepsize" [.05] "integer nused" [{nused}] "float maxdist" [0.4]
  "integer volumephotons" [{nphotons}]
Rotate 5 1 0 0
Camera "perspective" "float fov" [70]
WorldBegin
Translate 0 -0.5 3.5
Volume "homogeneous" "color sigma_a" [.05 .05 .05] "color sigma_s" [.1 .1 .1] "point p0" [-10 -10 -10] "point p1" [5 5 5]
AttributeBegin
LightSource "spot" "point from" [-3 0.72 0] "point to" [0 1.55 0] "color I" [15000 15000 15000] "float coneangle" [0.8]
LightSource "point" "point from" [0.1 1.35 -4] "color I" [4 4 4]
AttributeEnd
AttributeBegin
Material "glass" "float index" [1.3] "float Vn" [2.75] "color Kr" [0 0 0] "color Kt" [1 1 1]
Translate 0.1 1.35 0
Rotate 90 0 1 0
Scale 0.05 0.7 0.85
Shape "trianglemesh" "point P" [1 -1 -1  1 -1 1  -1 -1 1  -1 -1 -1  1 1 0  -1 1 0]
  "integer indices" [0 1 2  0 2 3  1 4 5  1 5 2  0 4 1  2 5 3  4 0 3  4 3 5]
AttributeEnd
Material "matte" "color Kd" [.001 .001 .001]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [5 -20 3  5 -20 -3  5 20 -3  5 20 3]
WorldEnd
"""


# The parameters of projectScene/scene.pbrt (the "crystal ball": a glass SPHERE in a homogeneous medium under a narrow spot
# light plus a point light, three matte quads, photonmap surface integrator with a caustic map and final gathering).
SPHERE_TEMPLATE = """# projectScene/scene.pbrt of the reference project: glass sphere (shapes/sphere.cpp) in a homogeneous medium
Film "image" "integer xresolution" [{xres}] "integer yresolution" [{yres}] "string filename" "{outfile}"
Sampler "lowdiscrepancy" "integer pixelsamples" [{spp}]
PixelFilter "{filt}"
SurfaceIntegrator "photonmap" "integer nused" [{surf_nused}] "bool finalgather" ["{finalgather}"] "integer finalgathersamples" [{fgsamples}]
  "float maxdist" [.25] "integer indirectphotons" [0] "integer causticphotons" [{caustic}]
VolumeIntegrator "photonvolume" "float stepsize" [.05] "integer nused" [{nused}] "float maxdist" [0.5]
  "integer volumephotons" [{nphotons}]
Rotate 5 1 0 0
Camera "perspective" "float fov" [70]
WorldBegin
Translate -1 -1 3.5
Volume "homogeneous" "color sigma_a" [.05 .05 .05] "color sigma_s" [.1 .1 .1] "point p0" [-10 0 -5] "point p1" [5 5 5]
AttributeBegin
LightSource "spot" "point from" [-3 5 0] "point to" [0 2 0] "color I" [2500 2500 2500] "float coneangle" [6]
LightSource "point" "point from" [0 2 -4] "color I" [8 8 8]
AttributeEnd
AttributeBegin
Material "glass" "color Kr" [{kr} {kr} {kr}] "color Kt" [1 1 1] "float index" [1.5] "float Vn" [{vn}]
Translate 0 2 0
{sphere_xform}Shape "sphere" "float radius" [.6]{sphere_params}
AttributeEnd
Material "matte" "color Kd" [.6 .6 .9]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-5 0 -5  5 0 -5  5 0 5  -5 0 5]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-5 0 3  5 0 3  5 10 3  -5 10 3]
Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [5 0 3  5 0 -3  5 10 -3  5 10 3]
WorldEnd
"""


def sphere_pbrt(nphotons=1000000, caustic=50000, finalgather=True, fgsamples=64, surf_nused=300, nused=300, xres=300, yres=300, spp=8,
                filt="gaussian", vn=0.0, kr=0.0, sphere_xform="", sphere_params="", outfile="scene.pfm"):
    return SPHERE_TEMPLATE.format(nphotons=nphotons, caustic=caustic, finalgather="true" if finalgather else "false", fgsamples=fgsamples,
                                  surf_nused=surf_nused, nused=nused, xres=xres, yres=yres, spp=spp, filt=filt, vn=vn, kr=kr,
                                  sphere_xform=sphere_xform, sphere_params=sphere_params, outfile=outfile)


def volumescene_pbrt(nphotons=5000, caustic=5000, finalgather=True, xres=300, yres=300, spp=1, filt="gaussian", outfile="volume.pfm"):
    return VOLUMESCENE_TEMPLATE.format(nphotons=nphotons, caustic=caustic, finalgather="true" if finalgather else "false", xres=xres,
                                       yres=yres, spp=spp, filt=filt, outfile=outfile)


def prism_pbrt(nphotons=5000000, caustic=1, nused=500, xres=512, yres=512, spp=32, filt="gaussian", outfile="pinkfloyd.pfm"):
    return PRISM_TEMPLATE.format(nphotons=nphotons, caustic=caustic, nused=nused, xres=xres, yres=yres, spp=spp, filt=filt, outfile=outfile)


def blob_density(n, seed=348, nblobs=8, floor=0.05):
    """Config-3 density: 8 seeded Gaussian blobs + floor, float32, index z*n*n + y*n + x
    (volumes/volumegrid.h:64).  Evaluated at voxel centres of [-1,1]^3."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-0.7, 0.7, size=(nblobs, 3)).astype(np.float32)
    widths = rng.uniform(0.15, 0.4, size=nblobs).astype(np.float32)
    amps = rng.uniform(0.5, 1.5, size=nblobs).astype(np.float32)
    c = ((np.arange(n, dtype=np.float32) + np.float32(0.5)) / np.float32(n)) * np.float32(2) - np.float32(1)
    out = np.full((n, n, n), np.float32(floor), dtype=np.float32)       # [z, y, x]
    for b in range(nblobs):
        dz = (c - centres[b, 2])[:, None, None]; dy = (c - centres[b, 1])[None, :, None]; dx = (c - centres[b, 0])[None, None, :]
        r2 = (dx * dx + dy * dy + dz * dz).astype(np.float32)
        out += (amps[b] * np.exp(-r2 / (np.float32(2) * widths[b] * widths[b]))).astype(np.float32)
    return np.ascontiguousarray(out.reshape(-1), dtype=np.float32)


def grid_volume_text(n, density, sigma_a=1.0, sigma_s=2.0, g=0.3):
    vals = " ".join("%.9g" % v for v in density)
    return ('Volume "volumegrid" "color sigma_a" [%g %g %g] "color sigma_s" [%g %g %g] "float g" [%g]\n'
            '  "point p0" [-1 -1 -1] "point p1" [1 1 1] "integer nx" [%d] "integer ny" [%d] "integer nz" [%d]\n'
            '  "float density" [%s]' % (sigma_a, sigma_a, sigma_a, sigma_s, sigma_s, sigma_s, g, n, n, n, vals))


def cornell_pbrt(volume_text, nphotons, xres=64, yres=64, stepsize=0.05, nused=50, maxdist=0.25, shoot_step=0.05,
                 outfile="cornell_vol.pfm", title="synthetic Cornell box (SURVEY.md Appendix C)"):
    return CORNELL_TEMPLATE.format(title=title, outfile=outfile, xres=xres, yres=yres, shoot_step=shoot_step,
                                   stepsize=stepsize, nused=nused, maxdist=maxdist, nphotons=nphotons, volume=volume_text)


# the Cornell box under the reference's other two volume integrators ("single": integrators/single.cpp,
# "emission": integrators/emission.cpp).  No photon maps: the surface integrator is the plain direct-lighting one.
VOLINT_MEDIA = {
    # emitting homogeneous medium, point + spot light (the light choice of single.cpp:117-119 matters)
    "volint_homog": ('Volume "homogeneous" "color sigma_a" [.3 .3 .3] "color sigma_s" [.15 .2 .25] "color Le" [.4 .2 .1] "float g" [0.4]\n'
                     '  "point p0" [-1 -1 -1] "point p1" [1 1 1]', True, 0.05),
    # optically thick homogeneous medium: the cumulative transmittance drops below 1e-3 and the Russian roulette runs
    "volint_dense": ('Volume "homogeneous" "color sigma_a" [2 2 2] "color sigma_s" [3 2.5 2] "color Le" [.05 .1 .2] "float g" [0]\n'
                     '  "point p0" [-1 -1 -1] "point p1" [1 1 1]', False, 0.05),
}
VOLINT_SPOT = 'LightSource "spot" "point from" [0.6 0.7 -0.6] "point to" [-0.2 -0.5 0.2] "color I" [30 25 20] "float coneangle" [40] "float conedeltaangle" [10]'


def volint_pbrt(kind, volume_text, stepsize=0.05, second_light=False, xres=64, yres=64, outfile="volint.pfm"):
    """Cornell box rendered with VolumeIntegrator `kind` ("single" | "emission")."""
    text = cornell_pbrt(volume_text, 0, xres=xres, yres=yres, stepsize=stepsize, outfile=outfile,
                        title="synthetic Cornell box, VolumeIntegrator \"%s\"" % kind)
    head, rest = text.split('SurfaceIntegrator "photonmap"', 1)
    rest = rest.split("LookAt", 1)[1]
    text = (head + 'SurfaceIntegrator "directlighting"\nVolumeIntegrator "%s" "float stepsize" [%g]\nLookAt' % (kind, stepsize) + rest)
    if second_light:
        text = text.replace('Material "matte" "color Kd" [.6 .6 .6]', VOLINT_SPOT + '\nMaterial "matte" "color Kd" [.6 .6 .6]', 1)
    return text


def volint_e2e_pbrt(kind, volume_text, stepsize=0.05, xres=72, yres=72, spp=4, outfile="volint_e2e.pfm"):
    """End-to-end scene for the "single" / "emission" drop-in: the all-maps Cornell box (glass wedge => specular bounces whose
    rays reach the volume integrator one at a time) under the direct-lighting surface integrator, two lights."""
    text = cornell_surf_pbrt(xres=xres, yres=yres, stepsize=stepsize, volume_text=volume_text, outfile=outfile)
    head, rest = text.split('SurfaceIntegrator "photonmap"', 1)
    rest = rest.split("LookAt", 1)[1]
    text = head + 'SurfaceIntegrator "directlighting"\nVolumeIntegrator "%s" "float stepsize" [%g]\nLookAt' % (kind, stepsize) + rest
    text = text.replace('"integer pixelsamples" [1]', '"integer pixelsamples" [%d]' % spp)
    return text.replace('Material "matte" "color Kd" [.6 .6 .6]', VOLINT_SPOT + '\nMaterial "matte" "color Kd" [.6 .6 .6]', 1)


def volint_offpath_pbrt(stepsize=0.05):
    """An "emission" scene whose surfaces and lights are OFF the device path (a disk shape that is an area light): the drop-in
    exports the medium alone for it (host/pv_export.inl, medium_only)."""
    text = volint_pbrt("emission", VOLINT_MEDIA["volint_homog"][0], stepsize=stepsize)
    return text.replace("WorldEnd", 'AttributeBegin\nAreaLightSource "diffuse" "color L" [5 5 5]\nTranslate 0 0.9 0\nRotate 90 1 0 0\n'
                                    'Shape "disk" "float radius" [0.3]\nAttributeEnd\nWorldEnd')


def aggregate_volumes(n=32):
    """Two overlapping Volume statements => the reference wraps them in an AggregateVolume (core/api.cpp:1200-1205,
    core/volume.cpp:178-261): an emitting, forward-scattering homogeneous slab in the left part of the box and the emitting blob
    grid squeezed into the right part; they share the slice -0.2 < x < 0.3."""
    homog = ('Volume "homogeneous" "color sigma_a" [.3 .3 .3] "color sigma_s" [.15 .2 .25] "color Le" [.4 .2 .1] "float g" [0.4]\n'
             '  "point p0" [-1 -1 -1] "point p1" [0.3 1 1]')
    grid = volint_grid_volume(n).replace('"point p0" [-1 -1 -1]', '"point p0" [-0.2 -1 -1]')
    return homog + "\n" + grid


AREA_QUAD = ('AttributeBegin\nAreaLightSource "diffuse" "color L" [6 5 4]\n'
             'Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-0.4 0.9 -0.4  0.4 0.9 -0.4  0.4 0.9 0.4  -0.4 0.9 0.4]\nAttributeEnd')


def volint_area_pbrt(kind="single", stepsize=0.05):
    """The Cornell box with a DiffuseAreaLight (a downward-facing quad under the ceiling = a ShapeSet of two triangles) next to the
    point light: groundwork for area lights on the device path."""
    return volint_pbrt(kind, VOLINT_MEDIA["volint_homog"][0], stepsize=stepsize).replace("WorldEnd", AREA_QUAD + "\nWorldEnd")


def volint_grid_volume(n=32):
    """Emitting, forward-scattering density grid (the config-3 blobs at n^3)."""
    return grid_volume_text(n, blob_density(n)).replace('"float g"', '"color Le" [.3 .3 .1] "float g"')


def cornell_surf_pbrt(nphotons=3000, caustic=1500, indirect=4000, finalgather=True, fgsamples=8, surf_nused=50, surf_maxdist=0.25,
                      xres=64, yres=64, stepsize=0.05, nused=50, maxdist=0.25, shoot_step=0.05, vn=0.0, volume_text=None,
                      outfile="cornell_surf.pfm"):
    return CORNELL_SURF_TEMPLATE.format(outfile=outfile, xres=xres, yres=yres, surf_nused=surf_nused, surf_maxdist=surf_maxdist,
                                        finalgather="true" if finalgather else "false", fgsamples=fgsamples, indirect=indirect,
                                        caustic=caustic, shoot_step=shoot_step, stepsize=stepsize, nused=nused, maxdist=maxdist,
                                        nphotons=nphotons, vn=vn, volume=volume_text or HOMOG_VOLUME)


def camera_rays(xres, yres, fov_deg=40.0, eye=(0.0, 0.0, -3.4), look=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0),
                u_scatter=0.5, y0=0, y1=None):
    """Pinhole rays through pixel centres (fov spans the shorter image axis, like
    cameras/perspective.cpp).  Rows y0..y1 only, row-major."""
    from .sceneio import make_rays
    y1 = yres if y1 is None else y1
    eye = np.asarray(eye, np.float64); look = np.asarray(look, np.float64); upv = np.asarray(up, np.float64)
    fwd = look - eye; fwd /= np.linalg.norm(fwd)
    right = np.cross(upv / np.linalg.norm(upv), fwd); right /= np.linalg.norm(right)
    upn = np.cross(fwd, right)
    aspect = xres / yres
    t = np.tan(np.radians(fov_deg) / 2)
    sx, sy = (aspect, 1.0) if aspect > 1 else (1.0, 1.0 / aspect)
    px = ((np.arange(xres) + 0.5) / xres * 2 - 1) * sx * t
    py = (1 - (np.arange(y0, y1) + 0.5) / yres * 2) * sy * t
    X, Y = np.meshgrid(px, py)
    d = X[..., None] * right + Y[..., None] * upn + fwd
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    d = d.reshape(-1, 3).astype(np.float32)
    o = np.broadcast_to(eye.astype(np.float32), d.shape)
    return make_rays(o, d, 0.0, np.inf, 0.0, u_scatter)


def synthetic_photons(n, seed=0x5EED, lo=-1.0, hi=1.0):
    """Kernel-isolated photon set: positions uniform in the medium box,
    wi uniform on the sphere, alpha a gently coloured spectrum of total weight ~1/n."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    z = rng.uniform(-1, 1, size=n); phi = rng.uniform(0, 2 * np.pi, size=n)
    r = np.sqrt(np.maximum(0, 1 - z * z))
    wi = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1).astype(np.float32)
    base = (1.0 + 0.01 * np.arange(30, dtype=np.float32))[None, :]
    scale = rng.uniform(0.5, 1.5, size=(n, 1)).astype(np.float32)
    alpha = (base * scale / np.float32(n)).astype(np.float32)
    return pos, wi, alpha
