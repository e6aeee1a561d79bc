"""cindm_b200.data.NBodyDataset (numpy reader for --initialization_mode 1/2) against the reference's own class
(data/nbody_dataset.py, loaded unmodified with torch_geometric / cindm.filepath stubbed) on a synthetic file."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

from cindm_b200.data import NBodyDataset, first_batch_1d, get_item_1d

REF = "/root/reference/data/nbody_dataset.py"


def make_file(tmp_path, n_bodies, n_simu, total_name):
    rng = np.random.default_rng(7)
    data = rng.uniform(0, 200, size=(n_simu, 1000, n_bodies, 4)).astype(np.float32)
    d = tmp_path / f"nbody-{n_bodies}"
    d.mkdir()
    np.save(d / f"trajectory_balls_{n_bodies}_simu_{total_name}_steps_1000.npy", data)
    return data


def load_reference_class():
    class Dataset:                      # torch_geometric.data.Dataset stand-in: indexing goes through get()
        def __init__(self, *a, **k):
            pass

    class Data(dict):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.__dict__.update(kw)

    saved = {k: sys.modules.get(k) for k in ("torch_geometric", "torch_geometric.data", "cindm", "cindm.filepath")}
    tg, tgd = types.ModuleType("torch_geometric"), types.ModuleType("torch_geometric.data")
    tgd.Dataset, tgd.Data = Dataset, Data
    tg.data = tgd
    cm, cf = types.ModuleType("cindm"), types.ModuleType("cindm.filepath")
    cf.NBODY_PATH = "/nonexistent"
    cm.filepath = cf
    sys.modules.update({"torch_geometric": tg, "torch_geometric.data": tgd, "cindm": cm, "cindm.filepath": cf})
    try:
        spec = importlib.util.spec_from_file_location("_ref_nbody_dataset", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.NBodyDataset


@pytest.mark.skipif(not os.path.isfile(REF), reason="reference tree not mounted")
@pytest.mark.parametrize("input_steps,output_steps", [(0, 24), (0, 44), (4, 20)])
def test_reader_matches_reference_class(tmp_path, input_steps, output_steps):
    make_file(tmp_path, 2, 3, 6000)
    ref_cls = load_reference_class()
    kw = dict(dataset="nbody-2", input_steps=input_steps, output_steps=output_steps, time_interval=4, is_y_diff=False,
              is_train=True, is_testdata=False)
    ref = ref_cls(dataset_path=str(tmp_path), **kw)
    ours = NBodyDataset(dataset_path=str(tmp_path), **kw)
    assert ours.time_stamps_effective == ref.time_stamps_effective
    assert ours.t_cushion_input == ref.t_cushion_input and ours.t_cushion_output == ref.t_cushion_output
    per_sim = ref.time_stamps_effective
    for idx in (0, 1, per_sim - 1, per_sim, 2 * per_sim + 17, 3 * per_sim - 1):
        a, b = ref.get(idx), ours.get(idx)
        assert (a.sim_id, a.time_id) == (b["sim_id"], b["time_id"])
        assert np.array_equal(a.y.numpy(), b["y"]) and np.array_equal(a.x.numpy(), b["x"])
    # the driver's first unshuffled batch in the diffusion layout (utils.get_item_1d :203-223)
    batch = [ref.get(i) for i in range(5)]
    y = torch.stack([d.y for d in batch]).reshape(-1, output_steps, 4)          # PyG batching: [B * n_bodies, steps, 4]
    want = (y.reshape(-1, 2, output_steps, 4) / 200.).permute(0, 2, 1, 3).flatten(-2, -1)
    got = first_batch_1d(ours, 5, "y")
    assert got.shape == (5, output_steps, 8) and torch.equal(got, want)


def test_split_sizes_and_clipping(tmp_path):
    data = make_file(tmp_path, 8, 6, 200)
    train = NBodyDataset(dataset="nbody-8", input_steps=0, output_steps=24, time_interval=4, dataset_path=str(tmp_path))
    assert train.n_simu == 6 and train.total_n_simu == 6                    # 180 of 200 in the reference; clipped to the file
    ev = NBodyDataset(dataset="nbody-8", input_steps=0, output_steps=24, time_interval=4, is_train=False,
                      dataset_path=str(tmp_path))
    assert ev.n_simu == 6
    full = NBodyDataset(dataset="nbody-2", input_steps=0, output_steps=24, time_interval=4, is_train=False,
                        data=np.zeros((6000, 800, 2, 4), dtype=np.float16))
    assert (full.total_n_simu, full.n_simu) == (6000, 100) and full.get(0)["sim_id"] == 5900   # last 100 simulations
    assert NBodyDataset(dataset="nbody-2", input_steps=0, output_steps=24, time_interval=4,
                        data=np.zeros((6000, 800, 2, 4), dtype=np.float16)).n_simu == 5800
    it = train.get(3)
    assert it["y"].shape == (8, 24, 4) and np.array_equal(it["y"][:, 0], data[0, 3 * 4 + 1])
    assert get_item_1d([it]).shape == (1, 24, 32)
    with pytest.raises(FileNotFoundError):
        NBodyDataset(dataset="nbody-4", dataset_path=str(tmp_path))
    with pytest.raises(IndexError):
        train.get(train.len())


def test_generator_draws_initial_states_in_the_reference_order(golden):
    """cindm_b200.data.nbody_simulation consumes Python's `random` exactly like the reference's generator (add_body per body,
    then the colour draws): the fixture was minted by executing the reference's own statements (oracle/make_golden_generator.py),
    so a seeded run starts from the states the reference would start from.  Flags and file name as in the reference."""
    import random
    from cindm_b200.data import nbody_simulation as gen
    g = golden("nbody_generator.npz")
    for name in (k for k in g.files if ":" not in k):
        seed, n = int(name.split("_")[0][4:]), int(name.split("_n")[1])
        random.seed(seed)
        states = gen.sample_initial_states(g[name].shape[0], n)
        assert np.array_equal(states, g[name]), name
        assert random.random() == float(g[name + ":next_random"])          # same number of draws consumed
        assert states[..., :2].min() >= 20 and states[..., :2].max() <= 180 and np.all(states[..., :2] == np.round(states[..., :2]))
    args = gen.build_parser().parse_args([])
    assert (args.n_bodies, args.n_simulations, args.vx, args.vy) == (2, 2, 100, 100)          # data/nbody_simulation.py:23-30
    assert gen.trajectory_filename("dataset/nbody_dataset", 2, 2, 100) == \
        "dataset/nbody_dataset/nbody-2/speed-100/trajectory_balls_2_simu_2_steps_1000.npy"   # :50
    with pytest.raises(ValueError):
        gen.main(["--n_bodies=9"])


@pytest.mark.gpu
def test_generator_end_to_end_on_the_cuda_rollout(tmp_path, golden):
    """The generator CLI: file in the reference's layout, frame 0 = the drawn states, every later frame bit-identical to the
    C oracle's rollout, and readable by the dataset reader the sampling driver uses."""
    from cindm_b200.data import NBodyDataset, nbody_simulation as gen
    from oracle import nbody_ref
    path = gen.main(["--n_bodies=4", "--n_simulations=5", "--seed=123", f"--dataset_root={tmp_path}"])
    assert path == str(tmp_path / "nbody-4" / "speed-100" / "trajectory_balls_4_simu_5_steps_1000.npy")
    data = np.load(path)
    assert data.shape == (5, 1000, 4, 4) and data.dtype == np.float64 and np.isfinite(data).all()
    states = golden("nbody_generator.npz")["seed123_n4"]
    assert np.array_equal(data[:, 0], states)
    assert np.array_equal(data, nbody_ref.rollout(states, 1000, 1))
    assert data[..., :2].min() > 15.0 and data[..., :2].max() < 185.0        # centres stay inside the walls up to one step's penetration
    ds = NBodyDataset(dataset="nbody-4", input_steps=4, output_steps=24, time_interval=4, is_y_diff=False, is_train=False,
                      is_testdata=False, data=data)
    assert len(ds) > 0
