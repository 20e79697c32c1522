"""GPU tests of the reference-facing training surface: train_loop end to end (checkpoints, resume, weights_dir, the
accumulation counter), Trainer.accumulate, trainer_from_config, hydra(float), multi-scale ground truth, call-level
corruption draws, the asynchronous step."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NAME = "resnet_color_1x6_bn_16x3x3_256x256_l1_relu"


def _pipeline_config(n_layers=2, batch=4, crop=32, images=12, **train):
    import blind_image_denoising_b200 as bf
    cfg = json.loads(json.dumps(bf.CONFIGS_DICT[NAME]))
    cfg["model"]["backbone"]["no_layers"] = n_layers
    cfg["train"] = {"epochs": 2, "total_steps": -1, "gpu_batches_per_step": 1, "checkpoints_to_keep": 2, "checkpoint_every": -1,
                    "visualization_every": 1,
                    "optimizer": {"type": "ADAM", "gradient_clipping_by_norm": 1.0,
                                  "schedule": {"type": "exponential_decay",
                                               "config": {"decay_rate": 0.9, "decay_steps": 100, "learning_rate": 0.002}}}}
    cfg["train"].update(train)
    cfg["loss"] = {"hinge": 0.5, "cutoff": 255.0, "mae_multiplier": 1.0, "ssim_multiplier": 0.0, "regularization": 0.01}
    cfg["dataset"] = {"batch_size": batch, "input_shape": [crop, crop, 3], "no_crops_per_image": 1, "seed": 0,
                      "additional_noise": [5, 40], "multiplicative_noise": [0.05, 0.1], "random_up_down": True,
                      "random_left_right": True,
                      "inputs": [{"synthetic": {"images": images, "height": crop + 8, "width": crop + 8, "seed": 7}}]}
    return cfg


def _metrics(d):
    return [json.loads(ln) for ln in open(os.path.join(d, "metrics.jsonl"))]


def test_train_loop_end_to_end_checkpoints_and_resume(native_lib, tmp_path):
    import bfcnn
    from blind_image_denoising_b200.train_loop import Checkpoint
    d = str(tmp_path / "run")
    cfg = _pipeline_config(epochs=2)
    assert bfcnn.train_loop(cfg, d) is None                                   # train_loop.py:40-43, returns nothing
    assert json.load(open(os.path.join(d, "pipeline.json")))["model"]["backbone"]["no_layers"] == 2    # :75-78
    assert os.path.isfile(os.path.join(d, "model_hydra", "variables", "variables.index"))              # :155-156
    # 12 images / batch 4 = 3 batches per epoch; the reference's counter sums k+1 = 2 micro-batches per update, so the
    # 6 batches of the two epochs make 3 updates... the counter runs ACROSS the epoch boundary (train_loop.py:404-437)
    ck = Checkpoint(None, d)
    step, epoch, variables = Checkpoint.read(ck.latest_checkpoint)
    assert (step, epoch) == (3, 2) and len(variables) == 3 + 5 * 2
    assert len(ck._existing()) == 2                                           # checkpoints_to_keep
    rec = _metrics(d)
    assert [r["step"] for r in rec[:3]] == [0, 1, 2] and rec[-1].get("final")
    assert all(np.isfinite(r["loss/total"]) and r["loss_denoiser/scale_0/mae"] > 0 for r in rec)
    assert rec[1]["training/learning_rate"] == pytest.approx(0.002 * 0.9 ** (1 / 100))
    first = Checkpoint.read(os.path.join(d, "ckpt-" + str(ck._existing()[0])))
    assert not np.array_equal(first[2][1], variables[1])                      # the weights moved
    # resume: the loop continues from (step, epoch) and trains one more epoch
    cfg3 = _pipeline_config(epochs=3)
    bfcnn.train_loop(cfg3, d)
    step2, epoch2, v2 = Checkpoint.read(Checkpoint(None, d).latest_checkpoint)
    assert epoch2 == 3 and step2 in (4, 5) and not np.array_equal(v2[1], variables[1])
    # the trained directory loads as a model (weights_dir / load_model read model_hydra or the checkpoints)
    d2 = str(tmp_path / "finetune")
    bfcnn.train_loop(_pipeline_config(epochs=0), d2, weights_dir=d)           # epochs 0: only loads + checkpoints
    s3, e3, v3 = Checkpoint.read(Checkpoint(None, d2).latest_checkpoint)
    assert (s3, e3) == (0, 0) and all(np.array_equal(a, b) for a, b in zip(v3, v2))
    with pytest.raises(ValueError, match="RMSPROP"):
        bad = _pipeline_config()
        del bad["train"]["optimizer"]["type"]
        bfcnn.train_loop(bad, str(tmp_path / "bad"))
    with pytest.raises(ValueError, match="gpu_batches_per_step"):
        bfcnn.train_loop(_pipeline_config(gpu_batches_per_step=0), str(tmp_path / "bad2"))


def test_train_loop_accumulation_counter_and_total_steps(native_lib, tmp_path):
    """k = 2: the reference applies after counter reaches 2, i.e. 3 micro-batches summed, scaled by 1/2; with
    exact_accumulation (extension) 2 micro-batches.  total_steps stops the loop."""
    import bfcnn
    from blind_image_denoising_b200.train_loop import Checkpoint
    cfg = _pipeline_config(images=24, epochs=-1, total_steps=2, gpu_batches_per_step=2)       # 6 batches per epoch
    d = str(tmp_path / "quirk")
    bfcnn.train_loop(cfg, d)
    s, e, _ = Checkpoint.read(Checkpoint(None, d).latest_checkpoint)
    assert s == 2 and e == 1                         # 2 updates = 6 micro-batches = exactly the first epoch
    cfg["train"]["exact_accumulation"] = True
    d = str(tmp_path / "exact")
    bfcnn.train_loop(cfg, d)
    s, e, _ = Checkpoint.read(Checkpoint(None, d).latest_checkpoint)
    assert s == 2 and e == 1                         # 2 updates = 4 micro-batches, still inside the first epoch
    # learning proceeds: the loss of a longer run goes down on the fixed synthetic set
    cfg = _pipeline_config(n_layers=2, images=8, batch=8, epochs=60, exact_accumulation=True, visualization_every=5)
    cfg["train"]["optimizer"]["schedule"]["config"]["learning_rate"] = 0.005
    cfg["dataset"]["multiplicative_noise"] = []
    d = str(tmp_path / "learn")
    bfcnn.train_loop(cfg, d, images=[np.tile(np.linspace(0, 255, 32, dtype=np.float32).astype(np.uint8)[None, :, None], (32, 1, 3)) for _ in range(8)])
    rec = _metrics(d)
    mae = [r["loss_denoiser/scale_0/mae"] for r in rec]
    assert np.mean(mae[-3:]) < 0.8 * np.mean(mae[:2]), mae


def test_accumulate_equals_one_big_batch_up_to_bn_statistics(native_lib):
    """Trainer.accumulate: k micro-batches then apply == the sum of the k gradients scaled 1/k (train_loop.py:418-434).
    With identical micro-batches the BN batch statistics agree as well, so the update equals a single-batch update."""
    import torch
    import blind_image_denoising_b200 as bf
    from blind_image_denoising_b200.training import Trainer
    from blind_image_denoising_b200 import _native
    arch = bf.Arch(no_layers=2)
    v = bf.synthetic_variables(arch, 0)
    opt = {"gpu_batches_per_step": 3, "schedule": {"type": "exponential_decay", "config": {"learning_rate": 1e-2, "decay_rate": 1.0, "decay_steps": 1}}}
    x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(4, 24, 24, 3), dtype=np.uint8)).cuda()
    nc = _native.NoiseCfg(10.0, 10.0, 0.0, 0.0, 0, 0, 0, 1)
    a, b = Trainer(arch, v, optimizer_config=opt), Trainer(arch, v, optimizer_config=opt)
    clean, noisy = a.prepare_data(x, nc, 1, 0)
    with pytest.raises(ValueError, match="accumulate"):
        a.apply_grads(None)                                    # nothing accumulated yet
    gs = []
    for k in range(3):
        _, _, _, g = a.train_step_single_gpu(clean, noisy, update_moving=False)
        gs.append(g.clone())
        done = a.accumulate(g)
        assert done == (k == 2)
    assert torch.allclose(a._accum, gs[0] + gs[1] + gs[2])
    a.apply_grads(None)
    _, _, _, g = b.train_step_single_gpu(clean, noisy, update_moving=False)
    assert float((a._accum / 3 - g).abs().max()) <= 1e-4 * float(g.abs().max())
    b.apply_grads(g)
    # Adam's first update is lr * g / (|g| + eps): where the gradient is (numerically) zero the sign of a 1e-6 run-to-run
    # difference (the BN statistics are reduced with atomics) decides the step, so compare where the gradient is not tiny
    from blind_image_denoising_b200.weights import flatten_variables, gather_trainables
    wa = gather_trainables(arch, flatten_variables(arch, a.get_weights()))
    wb = gather_trainables(arch, flatten_variables(arch, b.get_weights()))
    w0 = gather_trainables(arch, flatten_variables(arch, v))
    solid = (g.abs() > 1e-3 * g.abs().max()).cpu().numpy()
    assert solid.mean() > 0.5 and np.allclose(wa[solid], wb[solid], rtol=1e-5, atol=1e-6)
    assert np.allclose(np.abs(wa[solid] - w0[solid]), 1e-2, rtol=1e-2)      # the 1/k scale does not change a first Adam step
    assert not np.array_equal(a.get_weights()[1], v[1])
    # the accumulator restarts from zero for the next update
    _, _, _, g = a.train_step_single_gpu(clean, noisy, update_moving=False)
    a.accumulate(g)
    assert torch.allclose(a._accum, g)
    a.close(); b.close()


def test_trainer_from_config_and_async_step(native_lib):
    import torch
    import blind_image_denoising_b200 as bf
    from blind_image_denoising_b200 import _native
    cfg = _pipeline_config(n_layers=3, gpu_batches_per_step=2)
    cfg["loss"]["hinge"] = 2.0
    t = bf.trainer_from_config(cfg)
    assert t.arch == bf.Arch(no_layers=3) and t.gpu_batches_per_step == 2 and t.global_clipnorm == 1.0
    assert t.loss_cfg.hinge == 2.0 and t.loss_cfg.ssim_multiplier == 0.0 and t.schedule(0) == pytest.approx(0.002)
    bad = _pipeline_config()
    bad["train"]["optimizer"]["type"] = "RMSprop"
    with pytest.raises(ValueError, match="RMSPROP"):
        bf.trainer_from_config(bad)
    bad["train"]["optimizer"].update(type="Adam", amsgrad=True)
    with pytest.raises(ValueError, match="amsgrad"):
        bf.trainer_from_config(bad)
    # asynchronous step: same numbers as the synchronous one, fetched later
    x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(2, 20, 28, 3), dtype=np.uint8)).cuda()
    clean, noisy = t.prepare_data(x, _native.NoiseCfg(5.0, 40.0, 0.05, 0.1, 1, 1, 0, 1), 0, 0)
    total, ml, dl, g = t.train_step_single_gpu(clean, noisy, update_moving=False)
    g0 = g.clone()
    r = t.train_step_single_gpu(clean, noisy, update_moving=False, sync=False)
    assert r[0] is None and r[1] is None and r[2] is None
    total2, ml2, dl2 = t.last_losses()
    assert total2 == pytest.approx(total, rel=1e-6) and dl2["mae_loss"] == pytest.approx(dl["mae_loss"], rel=1e-6)
    assert float((r[3] - g0).abs().max()) <= 1e-4 * float(g0.abs().max())
    assert t.saved_activation("x", 3).shape == (2, 20, 28, 16) and float(t.saved_activation("t", 0).min()) >= 0.0
    with pytest.raises(Exception):
        t.saved_activation("t", 3)
    t.close()


def test_hydra_float_forward_matches_oracle(native_lib):
    """model_builder(...).hydra(x): float32 in, float32 out, moving statistics, no pow2 canvas (model.py:100-116)."""
    import bfcnn
    import blind_image_denoising_b200 as bf
    from oracle import bfcnn_oracle as O
    cfg = json.loads(json.dumps(bf.CONFIGS_DICT[NAME]["model"]))
    v = bf.synthetic_variables(bf.Arch(no_layers=6), 0)
    r = bfcnn.model_builder(cfg, variables=v)
    x = np.random.default_rng(0).uniform(-20, 280, size=(2, 37, 53, 3)).astype(np.float32)     # beyond [0,255]: the normaliser clips
    y = r.hydra(x)
    yref = O.hydra_forward(v, x.astype(np.float64))
    assert y.shape == x.shape and y.dtype == np.float32
    assert np.abs(y - yref).max() <= 0.02
    y2 = r.hydra([x])                                       # the reference passes a list (train_loop.py:248-256)
    assert np.array_equal(y, y2)
    with pytest.raises(RuntimeError, match="training=True"):
        r.hydra(x, training=True)
    r.hydra.close()


def test_multiscales_bit_exact(native_lib):
    import torch
    import blind_image_denoising_b200 as bf
    from blind_image_denoising_b200.training import Trainer
    from oracle import bfcnn_oracle as O
    t = Trainer(bf.Arch(no_layers=1), bf.synthetic_variables(bf.Arch(no_layers=1), 0))
    rng = np.random.default_rng(0)
    for shape in [(3, 64, 48, 3), (2, 37, 21, 3), (1, 2, 2, 3), (1, 1, 5, 3)]:
        x = rng.integers(0, 256, size=shape).astype(np.float32)           # rounded clean images (dataset.py:233-235)
        got = t.multiscales(torch.from_numpy(x).cuda(), 3)
        ref = O.multiscales(x, 3)
        assert len(got) == 4
        for a, b in zip(got, ref):
            assert tuple(a.shape) == b.shape and np.array_equal(a.cpu().numpy(), b)
    xf = (rng.uniform(-30, 300, size=(2, 16, 16, 3))).astype(np.float32)   # clip, and no rounding
    got = t.multiscales(torch.from_numpy(xf).cuda(), 1, clip_values=True, round_values=False)
    assert np.array_equal(got[1].cpu().numpy(), O.multiscales(xf, 1, round_values=False)[1])
    t.close()


def test_corrupt_draw_groups_bit_exact(native_lib):
    import torch
    import blind_image_denoising_b200 as bf
    from blind_image_denoising_b200.training import Trainer
    from blind_image_denoising_b200 import _native
    from oracle import corrupt_oracle as C
    t = Trainer(bf.Arch(no_layers=1), bf.synthetic_variables(bf.Arch(no_layers=1), 0))
    x = np.random.default_rng(2).integers(0, 256, size=(8, 12, 10, 3), dtype=np.uint8)
    for group, offset in ((4, 8), (2, 6), (-1, 5), (1, 3)):
        oc = C.NoiseConfig(draw_group=group)
        nc = _native.NoiseCfg(oc.additive_min, oc.additive_max, oc.multiplicative_min, oc.multiplicative_max, 1, 1, 0, 1, group)
        cr, nr = C.corrupt(x, 9, offset, oc)
        clean, noisy = t.prepare_data(torch.from_numpy(x).cuda(), nc, 9, offset)
        assert np.array_equal(clean.cpu().numpy(), cr)
        assert np.array_equal(noisy.cpu().numpy().view(np.uint32), nr.view(np.uint32)), (group, offset)
    t.close()


def test_image_pipeline_batches(native_lib):
    """ImagePipeline on the GPU: drop_remainder batching, every image used once per epoch, clean = rounded crop,
    noisy != clean, reshuffled every epoch, reproducible for a given epoch."""
    import torch
    import blind_image_denoising_b200 as bf
    import bfcnn
    from blind_image_denoising_b200.training import Trainer
    t = Trainer(bf.Arch(no_layers=1), bf.synthetic_variables(bf.Arch(no_layers=1), 0))
    imgs = [np.full((20, 20, 3), 10 * i, np.uint8) for i in range(11)]
    ds = bfcnn.dataset_builder({"batch_size": 4, "input_shape": [16, 16, 3], "no_crops_per_image": 2, "shuffle_buffer_images": 3,
                                "additional_noise": [5, 40], "random_left_right": True}, trainer=t, images=imgs)
    seen = []
    for clean, noisy in ds.training.epoch(0):
        assert clean.shape == (4, 16, 16, 3) and clean.dtype == torch.float32 and clean.is_cuda
        seen += [int(c[0, 0, 0]) for c in clean]
        assert float((clean - clean[:, :1, :1, :]).abs().max()) == 0.0
    assert len(seen) == (11 * 2 // 4) * 4 == 20
    assert all(seen.count(vv) <= 2 for vv in set(seen))
    again = [int(c[0, 0, 0]) for clean, _ in ds.training.epoch(0) for c in clean]
    other = [int(c[0, 0, 0]) for clean, _ in ds.training.epoch(1) for c in clean]
    assert again == seen and other != seen
    t.close()


def test_c_abi_allreduce_single_rank(native_lib):
    """bfcnn_allreduce_grads on a real ncclComm_t (one rank: the sum over ranks is the identity); bench.py's dp_check
    exercises it over N ranks.  Also: a NULL communicator is an argument error, not a crash."""
    import ctypes
    import torch
    import blind_image_denoising_b200 as bf
    from blind_image_denoising_b200 import _native
    from blind_image_denoising_b200.distributed import NcclCommunicator
    from blind_image_denoising_b200.training import Trainer
    arch = bf.Arch(no_layers=2)
    comm = NcclCommunicator(0, 1, NcclCommunicator.unique_id(), device=0)
    t = Trainer(arch, bf.synthetic_variables(arch, 0), nccl_comm=comm)
    g = torch.randn(arch.num_trainable(), device="cuda")
    g0 = g.clone()
    _native.check(t._lib.bfcnn_allreduce_grads(t.handle, g.data_ptr(), comm.comm, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert torch.equal(g, g0)
    w0 = t.get_weights()
    t.apply_grads(g)                      # all-reduce through the C ABI, then Adam
    assert not np.array_equal(t.get_weights()[1], w0[1])
    assert t._lib.bfcnn_allreduce_grads(t.handle, g.data_ptr(), None, None) == -1      # BFCNN_ERR_INVALID_ARGUMENT
    t.close()
    comm.close()
