"""CPU model of the device-side face-BVH build (rayito_b200/csrc/rt_build.cuh).

The CUDA build itself can only run on the GPU box (tests/test_gpu_build.py compares its trees byte for byte with
the host builder's).  What CAN be checked without a GPU is the method: that the three closed forms the kernels are
made of reproduce Bvh<T>::buildRange (Rayito_Stage7_QT/RAccel.h:290-374) exactly --

  1. std::partition's element order from ONE prefix sum of the predicate: with m elements satisfying it, the k-th
     "false" among the first m elements changes places with the k-th "true" from the right among the rest;
  2. the recursion's slot numbering from subtree sizes: children at base, base + 1, the left child's descendants
     from base + 2, the right child's from base + 2 * nL;
  3. in-order box unions with std::min / std::max tie rules as the minimum of (value with -0 == +0, position) keys,
     the sign of a winning zero carried along;

-- and that ranges of at most 32 elements finished by the serial recursion (literal libstdc++ partition loop) meet
the level-synchronous part seamlessly.  This file restates those steps in numpy, level by level like the kernels,
and compares the result node for node with the host builder (itself pinned to the compiled reference by
tests/test_host_parity.py / test_host_parallel.py)."""
import ctypes as C

import numpy as np
import pytest

SMALL = 32          # RT_BUILD_SMALL
F32MAX = np.float32(3.402823466e+38)


def _face_boxes(capi, scene, mesh_index=0):
    """Element boxes as Mesh::elementBBox computes them (RMesh.h:226-236): expand over the face's vertices in order,
    std::min / std::max keeping their first argument on ties."""
    d = scene.desc.contents
    mesh = (capi.RtMesh * d.num_meshes).from_address(C.cast(d.meshes, C.c_void_p).value)[mesh_index]
    verts = np.frombuffer(C.string_at(C.cast(d.vertices, C.c_void_p).value, 12 * d.num_vertices), np.float32).reshape(-1, 3)
    verts = verts[mesh.first_vertex:mesh.first_vertex + mesh.num_vertices]
    starts = np.frombuffer(C.string_at(C.cast(d.face_start, C.c_void_p).value, 4 * (d.num_faces + 1)), np.uint32)
    index = np.frombuffer(C.string_at(C.cast(d.vertex_index, C.c_void_p).value, 4 * d.num_indices), np.uint32)
    lo = np.full((mesh.num_faces, 3), F32MAX, np.float32)
    hi = np.full((mesh.num_faces, 3), -F32MAX, np.float32)
    first = starts[mesh.first_face:mesh.first_face + mesh.num_faces].astype(np.int64)
    count = (starts[mesh.first_face + 1:mesh.first_face + mesh.num_faces + 1] - starts[mesh.first_face:mesh.first_face + mesh.num_faces]).astype(np.int64)
    for k in range(int(count.max())):
        live = count > k
        p = verts[index[first[live] + k]]
        lo[live] = np.where(p < lo[live], p, lo[live])
        hi[live] = np.where(hi[live] < p, p, hi[live])
    nodes = np.frombuffer(C.string_at(C.cast(d.mesh_nodes, C.c_void_p).value + 32 * mesh.first_node, 32 * mesh.num_nodes),
                          np.uint32).reshape(-1, 8)
    return np.concatenate([lo, hi], axis=1), nodes


def _union_keys(values, want_min):
    """In-order union of a column with first-argument-wins ties: index of the winner = smallest (value with -0 == +0,
    position) for a minimum, largest value / smallest position for a maximum (rt_build.cuh min_key / max_key)."""
    canon = values + np.float32(0.0)        # -0 + 0 = +0: the two zeros compare equal in std::min / std::max
    order = np.lexsort((np.arange(len(values)), canon if want_min else -canon))
    return values[order[0]]                 # the winner's own bits, sign of zero included


def _union(boxes):
    out = np.empty(6, np.float32)
    for c in range(3):
        out[c] = _union_keys(boxes[:, c], True)
        out[3 + c] = _union_keys(boxes[:, 3 + c], False)
    return out


def _plan(box):
    ext = box[3:] - box[:3]             # float32 arithmetic, as buildRange (RAccel.h:305-326)
    if ext[0] > ext[1]:
        axis = 0 if ext[0] > ext[2] else 2
    else:
        axis = 1 if ext[1] > ext[2] else 2
    where = (box[3 + axis] + box[axis]) * np.float32(0.5)
    return axis, where


def _above(boxes, axis, where):
    return where < (boxes[:, 3 + axis] + boxes[:, axis]) * np.float32(0.5)


def _write(nodes, node, box, word, flags):
    nodes[node, :6] = box.view(np.uint32)
    nodes[node, 6] = word
    nodes[node, 7] = flags


def _serial(nodes, boxes, prims, job):
    """k_small: the reference's recursion, literally (explicit stack, libstdc++'s bidirectional partition loop)."""
    stack = [job]
    deepest = 0
    while stack:
        b, e, node, base, depth, box = stack.pop()
        deepest = max(deepest, depth)
        if e - b <= 1:
            _write(nodes, node, box, prims[b], 4)
            continue
        axis, where = _plan(box)
        _write(nodes, node, box, base, axis)
        first, last = b, e
        while True:
            while first != last and _above(boxes[first:first + 1], axis, where)[0]:
                first += 1
            if first == last:
                break
            last -= 1
            while first != last and not _above(boxes[last:last + 1], axis, where)[0]:
                last -= 1
            if first == last:
                break
            boxes[[first, last]] = boxes[[last, first]]
            prims[[first, last]] = prims[[last, first]]
            first += 1
        mid = first
        if mid <= b or mid >= e:
            mid = b + (e - b) // 2
        kids = []
        for side, (cb, ce) in enumerate(((b, mid), (mid, e))):
            lo = np.full(3, F32MAX, np.float32)
            hi = np.full(3, -F32MAX, np.float32)
            for i in range(cb, ce):
                lo = np.where(boxes[i, :3] < lo, boxes[i, :3], lo)
                hi = np.where(hi < boxes[i, 3:], boxes[i, 3:], hi)
            kids.append((cb, ce, base + side, base + 2 if side == 0 else base + 2 * (mid - b), depth + 1,
                         np.concatenate([lo, hi]).astype(np.float32)))
        stack.append(kids[1])
        stack.append(kids[0])
    return deepest


def build_model(face_boxes):
    n = len(face_boxes)
    boxes = face_boxes.copy()
    prims = np.arange(n, dtype=np.uint32)
    nodes = np.zeros((2 * n - 1, 8), np.uint32)
    level = [(0, n, 0, 1, 0, _union(boxes))]
    small = []
    if n <= SMALL:
        small, level = level, []
    while level:
        # one level: predicate of every element of every range, ONE prefix sum over the whole array
        pred = np.zeros(n + 1, np.int64)
        plans = []
        for (b, e, node, base, depth, box) in level:
            axis, where = _plan(box)
            _write(nodes, node, box, base, axis)
            pred[b:e] = _above(boxes[b:e], axis, where)
            plans.append((axis, where))
        scan = np.concatenate([[0], np.cumsum(pred[:-1])])        # exclusive, n + 1 entries
        nxt = []
        for (b, e, node, base, depth, box) in level:
            trues = int(scan[e] - scan[b])
            if trues == 0 or trues == e - b:
                mid = b + (e - b) // 2                              # nothing moved; cut in half
            else:
                mid = b + trues
                idx = np.arange(b, e)
                p = pred[b:e].astype(bool)
                left = idx[(idx < mid) & ~p]                        # k-th false from the left: k = (i - b) - (scan[i] - scan[b])
                right = idx[(idx >= mid) & p]                       # k-th true from the right: k = scan[e] - scan[i + 1]
                assert np.array_equal((left - b) - (scan[left] - scan[b]), np.arange(len(left)))
                assert np.array_equal(scan[e] - scan[right + 1], np.arange(len(right))[::-1])
                right = right[::-1]
                boxes[np.concatenate([left, right])] = boxes[np.concatenate([right, left])]
                prims[np.concatenate([left, right])] = prims[np.concatenate([right, left])]
            for side, (cb, ce) in enumerate(((b, mid), (mid, e))):
                child = (cb, ce, base + side, base + 2 if side == 0 else base + 2 * (mid - b), depth + 1, _union(boxes[cb:ce]))
                (nxt if ce - cb > SMALL else small).append(child)
        level = nxt
    deepest = 0
    for job in small:
        deepest = max(deepest, _serial(nodes, boxes, prims, job))
    return nodes, deepest


@pytest.mark.parametrize("grid", [(7, 5), (33, 1), (64, 48), (160, 96)])
def test_model_reproduces_the_host_tree_on_the_sphere(capi, grid):
    scene = capi.HostScene(capi.RECIPE_SYNTHETIC_MESH, None, grid)
    boxes, want = _face_boxes(capi, scene)
    got, depth = build_model(boxes)
    assert np.array_equal(got, want)
    assert depth == scene.depth(0)


def test_model_reproduces_the_host_tree_on_bumpy_and_the_wedge(capi, scene1_host, deepboth_host):
    boxes, want = _face_boxes(capi, scene1_host, 1)       # mesh 1 = bumpy.obj (mesh 0 is the six-quad cube)
    assert len(boxes) == 24576
    got, depth = build_model(boxes)
    assert np.array_equal(got, want) and depth == 20
    boxes, want = _face_boxes(capi, deepboth_host)
    got, depth = build_model(boxes)
    assert np.array_equal(got, want) and depth == deepboth_host.depth(0) >= 40


def test_zero_signs_follow_first_argument_wins():
    """A union over -0 and +0 keeps whichever came first, for the minimum and for the maximum."""
    z = np.array([0.0, -0.0, 0.0], np.float32)
    assert np.signbit(_union_keys(z[1:], True)) and not np.signbit(_union_keys(z, True))
    assert np.signbit(_union_keys(z[1:], False)) and not np.signbit(_union_keys(z, False))
    assert _union_keys(np.array([1.0, -0.0, 0.0, -2.0], np.float32), True) == np.float32(-2.0)
