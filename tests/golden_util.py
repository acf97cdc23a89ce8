"""Loads tests/golden/bvh_golden.npz (reference outputs, see tests/golden/gen_golden.py)."""
import os

import numpy as np

import scenes
from realtimeraytracing_b200 import synth

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bvh_golden.npz")
SCENE_NAMES = ["survey20k", "soup512", "soup65536", "two_mesh", "padded", "dupcodes", "grid"]


def load():
    return np.load(_PATH)


def scene(name):
    """(triangles, meshes, n) of a golden scene, regenerated from its seed."""
    if name == "survey20k":
        t, m = synth.survey_known_answer_scene()
        return t, m, t.size
    if name == "soup512":
        t, m, _ = scenes.soup(512)
        return t, m, t.size
    if name == "soup65536":
        t, m, _ = scenes.soup(65536)
        return t, m, t.size
    if name == "two_mesh":
        t, m = scenes.two_mesh_scene()
        return t, m, t.size
    if name == "padded":
        return scenes.padded_scene()
    if name == "dupcodes":
        t, m = scenes.duplicate_codes_scene()
        return t, m, t.size
    if name == "grid":
        t, m = synth.grid_mesh(40, 30)
        return t, m, t.size
    raise KeyError(name)
