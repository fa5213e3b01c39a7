"""CPU: the tree checker of tests/test_gpu_lbvh.py and Scene.prim_bounds() against the REFERENCE's own BVHs.  The exported
LinearBVHNode arrays of the seven golden scenes (built by the reference's SAH code, accelerators/bvh.cpp:196-577) must pass every
invariant the checker states -- permutation, every primitive in exactly one leaf, leaf and interior boxes equal to the unions of
Scene.prim_bounds() BIT FOR BIT (so prim_bounds restates Triangle::WorldBound / Sphere::WorldBound exactly, incl. the transformed
partial spheres), depth-first layout, depth <= 64 -- and a corrupted tree must not."""
import numpy as np
import pytest
from test_gpu_lbvh import ALL_SCENES, NODE, check_tree, brute_tree


@pytest.mark.parametrize("name", ALL_SCENES)
def test_reference_trees_pass_the_checker(golden, name):
    _, scene = golden(name)
    b = scene.prim_bounds()
    n_nodes, depth = check_tree(np.asarray(scene.nodes, np.uint8), np.arange(len(b)), b, 255, morton_ordered=False)
    assert n_nodes == scene.n_nodes and 1 <= depth <= 64


def test_checker_rejects_broken_trees(golden):
    _, scene = golden("cornell_homog")
    b = scene.prim_bounds()
    good = np.frombuffer(np.asarray(scene.nodes, np.uint8).tobytes(), dtype=NODE).copy()
    leaf = int(np.nonzero(good["n_primitives"] > 0)[0][0]); inner = int(np.nonzero(good["n_primitives"] == 0)[0][0])

    def broken(edit):
        nd = good.copy(); edit(nd)
        with pytest.raises(AssertionError):
            check_tree(np.frombuffer(nd.tobytes(), np.uint8), np.arange(len(b)), b, 255, morton_ordered=False)

    broken(lambda nd: nd["bounds"].__setitem__((leaf, 0), nd["bounds"][leaf, 0] - 1.0))        # a leaf box that is not the union
    broken(lambda nd: nd["bounds"].__setitem__((inner, 3), nd["bounds"][inner, 3] + 1.0))      # an interior box that is not the union
    broken(lambda nd: nd["offset"].__setitem__(inner, nd["offset"][inner] + 1))                # second child not where the first subtree ends
    broken(lambda nd: nd["n_primitives"].__setitem__(leaf, nd["n_primitives"][leaf] + 1))      # a primitive in two leaves / out of range
    broken(lambda nd: nd["axis"].__setitem__(inner, 3))


def test_brute_tree_is_a_valid_tree():
    rs = np.random.RandomState(3)
    v = rs.uniform(-1, 1, (1000, 3, 3)).astype(np.float32)
    b = np.concatenate([v.min(axis=1), v.max(axis=1)], axis=1).astype(np.float32)
    n_nodes, depth = check_tree(brute_tree(b), np.arange(len(b)), b, 255, morton_ordered=False)
    assert n_nodes == 7 and depth == 3                       # 1000 primitives -> 4 leaves of 250 under a balanced hierarchy
