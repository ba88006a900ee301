"""Generate tests/golden/augment_reference.npz (run in THIS container only).

Executes the reference's OWN ``point_removal``, ``random_noise`` and ``rotate_points``
(/root/reference/augmentation.py:54-122) in place -- their source is pulled out of the reference file with ``ast`` at
run time (the module itself cannot be imported: it needs laspy and torch_geometric) and is never copied into this
repo -- on seeded inputs with seeded ``random`` / ``numpy.random`` generators, and stores inputs, outputs and the
random draws the functions consumed (recovered by replaying the generators' call sequence).  The fixture pins
oracle/augment_ref.apply_augmentation.
"""
import ast
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
REF_FILE = "/root/reference/augmentation.py"


def reference_functions():
    tree = ast.parse(open(REF_FILE).read())
    want = ("rotate_points", "point_removal", "random_noise")
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    ns = {"np": np, "random": random}
    exec(compile(ast.Module(body=fns, type_ignores=[]), REF_FILE, "exec"), ns)
    return [ns[k] for k in want]


def replay_draws(n, dim, s):
    """The draws the three functions make, in their call order, from generators seeded like the run below."""
    random.seed(s)
    np.random.seed(s)
    idx = list(range(n))
    random.shuffle(idx)
    keep = np.random.choice(idx, random.randint(round(len(idx) * 0.9), len(idx)), replace=False)
    m = len(keep)
    sd = np.random.uniform(0.01, 0.025)
    add = np.random.uniform(0, 1) >= 0.5
    noise_c = np.random.normal(0, sd, size=(m, 3))
    noise_x = np.random.normal(0, sd, size=(m, dim))
    use = np.random.choice(m, random.randint(0, round(m * 0.1)), replace=False)
    angle = np.random.uniform(-180, 180)
    return keep, noise_c, noise_x, add, use, angle


def main():
    rotate_points, point_removal, random_noise = reference_functions()
    out = {}
    cases = [(101, 1, 7), (1000, 1, 8), (3000, 1, 9), (513, 2, 10), (2500, 1, 11)]
    for ci, (n, dim, s) in enumerate(cases):
        rng = np.random.default_rng(1000 + s)
        coords = rng.normal(size=(n, 3)) * np.array([4.0, 4.0, 8.0])
        x = rng.uniform(0, 20, size=(n, dim))
        random.seed(s)
        np.random.seed(s)
        c, xx = point_removal(coords.copy(), x.copy())
        c, xx = random_noise(c, dim, xx)
        c = rotate_points(c)
        keep, noise_c, noise_x, add, use, angle = replay_draws(n, dim, s)
        out.update({f"c{ci}_coords": coords, f"c{ci}_x": x, f"c{ci}_out_coords": c, f"c{ci}_out_x": xx,
                    f"c{ci}_keep": keep, f"c{ci}_noise_c": noise_c, f"c{ci}_noise_x": noise_x,
                    f"c{ci}_add": np.array(add), f"c{ci}_use": use, f"c{ci}_angle": np.array(angle)})
    out["num_cases"] = np.array(len(cases))
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "augment_reference.npz"), **out)
    print("wrote augment_reference.npz:", len(cases), "cases")


if __name__ == "__main__":
    main()
