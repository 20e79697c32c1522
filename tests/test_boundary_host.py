"""CPU checks of the reference-facing builders (bfcnn/__init__.py:129-143 names): optimizer / schedule builders against
the Keras closed forms, model_builder's BuilderResults, config rejections, (step, epoch, model) checkpoints, the image
pipeline's host side, and the oracle restatements added for them.  No compute call reaches the GPU here."""
import json
import math
import os

import numpy as np
import pytest

import bfcnn
import blind_image_denoising_b200 as bf
from blind_image_denoising_b200 import _native
from blind_image_denoising_b200.arch import Arch, arch_from_config
from blind_image_denoising_b200.optimizer import optimizer_config_check
from blind_image_denoising_b200.train_loop import Checkpoint, create_checkpoint, load_config
from blind_image_denoising_b200.training import ImagePipeline, noise_cfg_from_config


def test_reference_names_are_importable_from_bfcnn():
    # reference bfcnn/__init__.py:129-143 (the hot-path subset of __all__) + the builders SURVEY 8b lists
    for name in ("models", "configs", "train_loop", "load_model", "load_image", "model_builder", "schedule_builder",
                 "optimizer_builder", "load_denoiser_model", "load_default_denoiser", "dataset_builder",
                 "loss_function_builder", "CONFIGS_DICT"):
        assert hasattr(bfcnn, name), name
    import inspect
    sig = inspect.signature(bfcnn.train_loop)
    assert list(sig.parameters)[:3] == ["pipeline_config_path", "checkpoint_directory", "weights_dir"]   # train_loop.py:40-43
    assert sig.parameters["weights_dir"].default is None


def test_schedule_builder_matches_keras_closed_forms():
    # keras ExponentialDecay: lr * rate ** (step / steps)
    s = bfcnn.schedule_builder({"type": "exponential_decay", "config": {"learning_rate": 1e-3, "decay_rate": 0.9, "decay_steps": 40000}})
    assert s(0) == pytest.approx(1e-3) and s(40000) == pytest.approx(9e-4) and s(20000) == pytest.approx(1e-3 * 0.9 ** 0.5)
    # keras CosineDecay(alpha default 0.0001 in optimizer.py:131): ((1-a) * 0.5 (1 + cos(pi p)) + a) * lr, p clipped at 1
    c = bfcnn.schedule_builder({"type": "cosine_decay", "config": {"learning_rate": 2e-3, "decay_steps": 1000}})
    assert c(0) == pytest.approx(2e-3)
    assert c(500) == pytest.approx(2e-3 * ((1 - 1e-4) * 0.5 + 1e-4))
    assert c(1000) == pytest.approx(2e-3 * 1e-4) and c(5000) == pytest.approx(2e-3 * 1e-4)
    # keras CosineDecayRestarts(t_mul 2, m_mul 0.9, alpha 0.001): periods 100, 200, 400, ...; amplitude x 0.9 per restart
    r = bfcnn.schedule_builder({"type": "cosine_decay_restarts", "config": {"learning_rate": 1.0, "decay_steps": 100}})
    assert r(0) == pytest.approx(1.0)
    assert r(50) == pytest.approx((1 - 1e-3) * 0.5 + 1e-3)
    assert r(100) == pytest.approx((1 - 1e-3) * 0.9 + 1e-3)          # start of the second period
    assert r(200) == pytest.approx((1 - 1e-3) * 0.9 * 0.5 + 1e-3)     # middle of the second period (length 200)
    assert r(300) == pytest.approx((1 - 1e-3) * 0.81 + 1e-3)          # start of the third
    r1 = bfcnn.schedule_builder({"type": "cosine_decay_restarts", "config": {"learning_rate": 1.0, "decay_steps": 10, "t_mul": 1.0, "m_mul": 0.5}})
    assert r1(25) == pytest.approx((1 - 1e-3) * 0.25 * 0.5 * (1 + math.cos(math.pi * 0.5)) + 1e-3)
    # argument checks of optimizer.py:83-103
    with pytest.raises(ValueError, match="config must be a dictionary"):
        bfcnn.schedule_builder(None)
    with pytest.raises(ValueError, match="schedule_type cannot be None"):
        bfcnn.schedule_builder({})
    with pytest.raises(ValueError, match="schedule_type must be a string"):
        bfcnn.schedule_builder({"type": 3})
    with pytest.raises(ValueError, match="don't know how to handle"):
        bfcnn.schedule_builder({"type": "linear", "config": {"learning_rate": 1.0}})


def test_optimizer_builder_only_adam_and_says_so():
    sched = {"type": "exponential_decay", "config": {"learning_rate": 1e-3, "decay_rate": 0.9, "decay_steps": 100}}
    opt, lr = bfcnn.optimizer_builder({"type": "ADAM", "gradient_clipping_by_norm": 1.0, "schedule": sched})   # the in-tree config
    assert opt.name == "Adam" and opt.global_clipnorm == 1.0 and (opt.beta_1, opt.beta_2, opt.epsilon) == (0.9, 0.999, 1e-7)
    assert opt.iterations == 0 and opt.learning_rate == pytest.approx(1e-3) and lr(100) == pytest.approx(9e-4)
    # the reference's default type is RMSprop (optimizer.py:165): a config without a type must not train with Adam silently
    with pytest.raises(ValueError, match="RMSPROP"):
        bfcnn.optimizer_builder({"schedule": sched})
    with pytest.raises(ValueError, match="ADADELTA"):
        bfcnn.optimizer_builder({"type": "Adadelta", "schedule": sched})
    with pytest.raises(ValueError, match="don't know how to handle optimizer_type"):
        bfcnn.optimizer_builder({"type": "sgd", "schedule": sched})
    for key in ("amsgrad", "gradient_clipping_by_value", "gradient_clipping_by_norm_local"):
        with pytest.raises(ValueError, match=key.split("_")[0]):
            optimizer_config_check({"type": "Adam", key: 1.0})
    with pytest.raises(ValueError, match="config must be a dictionary"):
        bfcnn.optimizer_builder([])


def test_deep_supervision_schedules():
    from blind_image_denoising_b200.optimizer import deep_supervision_schedule_builder as dsb
    assert dsb({"type": "linear_low_to_high"}, 1)(0.3).tolist() == [1.0]       # single-output resnet: always [1]
    d = dsb({"type": "linear_low_to_high"}, 3)
    assert np.allclose(d(0.0), [1 / 6, 2 / 6, 3 / 6]) and np.allclose(d(1.0), [3 / 6, 2 / 6, 1 / 6])
    assert np.allclose(dsb({"type": "constant_equal"}, 4)(0.5), [0.25] * 4)
    with pytest.raises(ValueError):
        dsb({"type": "nope"}, 2)
    with pytest.raises(ValueError):
        dsb({"type": "constant_equal"}, 0)


def test_model_builder_results():
    cfg = json.loads(json.dumps(bf.CONFIGS_DICT["resnet_color_1x6_bn_16x3x3_256x256_l1_relu"]["model"]))
    r = bfcnn.model_builder(cfg)
    assert r._fields == ("backbone", "normalizer", "denormalizer", "denoiser", "hydra", "options")   # model.py:25-34
    arch = Arch(no_layers=6)
    assert [v.shape for v in r.hydra.variables] == [tuple(s) for s in arch.variable_shapes()]
    assert len(r.hydra.trainable_variables) == 1 + 3 * 6 + 2 and len(r.hydra.non_trainable_variables) == 2 * 6
    assert r.backbone.count_params() + r.denoiser.count_params() == arch.num_weights() == r.hydra.count_params()
    assert r.normalizer.variables == [] and r.denormalizer.variables == [] and len(r.hydra.outputs) == 1
    # Keras initial state: gamma 1, moving mean 0, moving variance 1; kernels glorot-scaled
    v = r.hydra.variables
    assert np.all(v[3] == 1) and np.all(v[4] == 0) and np.all(v[5] == 1)
    assert 0.5 < v[1].std() / math.sqrt(2.0 / (9 * 16 + 9 * 16)) < 1.5
    w = [x + 1 for x in r.hydra.get_weights()]
    r.hydra.set_weights(w)
    assert np.array_equal(r.hydra.variables[0], w[0])
    with pytest.raises(ValueError):
        r.hydra.set_weights(w[:-1])
    with pytest.raises(RuntimeError, match="no stand-alone kernel"):
        r.normalizer(np.zeros((1, 4, 4, 3)))
    lines = []
    r.hydra.summary(print_fn=lines.append)
    assert any("28784" in ln for ln in lines)
    with pytest.raises(KeyError):
        bfcnn.model_builder({"backbone": cfg["backbone"]})


@pytest.mark.parametrize("section,key,value", [
    ("backbone", "activation", "leaky_relu"), ("backbone", "base_activation", "relu"), ("backbone", "kernel_regularizer", "l2"),
    ("backbone", "dropout_rate", 0.1), ("backbone", "add_gradient_dropout", True), ("backbone", "block_groups", [1, 2]),
    ("backbone", "block_depthwise", [-1, 4]), ("backbone", "block_activation", ["linear", "relu"]),
    ("backbone", "block_regularizer", ["l1", "l2"]), ("backbone", "selector_params", {"a": 1}),
    ("backbone", "value_range", [0, 1]), ("backbone", "add_gates", True), ("backbone", "use_bias", True),
    ("denoiser", "activation", "relu"), ("denoiser", "use_bn", True), ("denoiser", "use_ln", True),
    ("denoiser", "kernel_regularizer", "l1"),
])
def test_arch_from_config_rejects_what_the_kernels_do_not_compute(section, key, value):
    cfg = json.loads(json.dumps(bf.CONFIGS_DICT["resnet_color_1x6_bn_16x3x3_256x256_l1_relu"]))
    assert arch_from_config(cfg) == Arch(no_layers=6)
    cfg["model"][section][key] = value
    with pytest.raises(ValueError, match="bias" if key == "use_bias" else key):
        arch_from_config(cfg)


def test_arch_from_config_accepts_the_reference_spellings():
    cfg = json.loads(json.dumps(bf.CONFIGS_DICT["resnet_color_1x6_bn_16x3x3_256x256_l1_relu"]))
    bb = cfg["model"]["backbone"]
    bb.update({"block_activation": ["relu", "linear"], "block_regularizer": ["l1", "l1"], "block_groups": [1, 1],
               "block_depthwise": [-1, -1], "activation": " ReLU ", "dropout_rate": -1})
    assert arch_from_config(cfg) == Arch(no_layers=6)
    assert "bf16" not in _native.PRECISIONS     # the tensor-core arm computes in fp16, it is not called bf16


def test_noise_cfg_from_config():
    c = noise_cfg_from_config({"additional_noise": [5, 40, 20], "multiplicative_noise": [0.1, 0.05], "random_up_down": True,
                               "round_values": False, "no_crops_per_image": 4})
    assert (c.additive_min, c.additive_max) == (5.0, 40.0)
    assert c.multiplicative_min == pytest.approx(0.05) and c.multiplicative_max == pytest.approx(0.1)
    assert (c.random_left_right, c.random_up_down, c.subsample) == (0, 1, 0)
    assert c.round_values == 1          # dataset.py:228 rounds whatever the config says
    assert c.draw_group == 4            # the crops of one image share the call-level draws (dataset.py:276-297)
    e = noise_cfg_from_config({})
    assert (e.additive_max, e.multiplicative_max, e.draw_group) == (0.0, 0.0, 1)
    with pytest.raises(ValueError):
        noise_cfg_from_config({"use_jpeg_noise": True})


class _FakeModel:
    def __init__(self, n=3):
        self.w = [np.full((2, 2), float(i), np.float32) for i in range(n)]

    def get_weights(self):
        return [x.copy() for x in self.w]

    def set_weights(self, w):
        self.w = [np.asarray(x, np.float32) for x in w]


def test_checkpoint_manager_semantics(tmp_path):
    """train_loop.py:149-181: ckpt-<n> files, `checkpoint` state file, max_to_keep, restore latest."""
    m = _FakeModel()
    ck = create_checkpoint(model=m, path=None, directory=tmp_path, max_to_keep=2)
    assert ck.latest_checkpoint is None
    paths = []
    for step in (0, 10, 20):
        ck.step, ck.epoch = step, step // 10
        m.w[0][:] = step
        paths.append(ck.save())
    assert [os.path.basename(p) for p in paths] == ["ckpt-1", "ckpt-2", "ckpt-3"]
    names = sorted(os.listdir(tmp_path))
    assert names == ["checkpoint", "ckpt-2.data-00000-of-00001", "ckpt-2.index", "ckpt-3.data-00000-of-00001", "ckpt-3.index"]
    state = open(tmp_path / "checkpoint").read()
    assert 'model_checkpoint_path: "ckpt-3"' in state and state.count("all_model_checkpoint_paths") == 2
    m2 = _FakeModel()
    ck2 = create_checkpoint(model=m2, path=tmp_path, directory=tmp_path)
    assert (ck2.step, ck2.epoch) == (20, 2) and np.all(m2.w[0] == 20) and np.all(m2.w[2] == 2)
    assert os.path.basename(ck2.save()) == "ckpt-4"              # numbering goes on after a restart
    step, epoch, variables = Checkpoint.read(str(tmp_path / "ckpt-3"))
    assert (step, epoch, len(variables)) == (20, 2, 3)
    # a model checkpointed through HydraModel round-trips its 33 variables in integer order ("10" < "2" on disk)
    r = bfcnn.model_builder(json.loads(json.dumps(bf.CONFIGS_DICT["resnet_color_1x6_bn_16x3x3_256x256_l1_relu"]["model"])))
    ck3 = create_checkpoint(model=r.hydra, directory=tmp_path / "h")
    ck3.step = 7
    p = ck3.save()
    s2, _, v2 = Checkpoint.read(p)
    assert s2 == 7 and all(np.array_equal(a, b) for a, b in zip(v2, r.hydra.variables))


def test_load_config_errors(tmp_path):
    assert load_config({"a": 1}) == {"a": 1}
    with pytest.raises(ValueError, match="config should not be empty"):
        load_config(None)
    with pytest.raises(ValueError, match="is not valid"):
        load_config(str(tmp_path / "missing.json"))
    (tmp_path / "c.json").write_text('{"b": 2}')
    assert load_config(tmp_path / "c.json") == {"b": 2}


def test_image_pipeline_host_side(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(0)
    for i in range(5):
        Image.fromarray(rng.integers(0, 256, size=(40 + i, 50, 3), dtype=np.uint8)).save(tmp_path / f"im{i}.png")
    Image.fromarray(rng.integers(0, 256, size=(8, 8, 3), dtype=np.uint8)).save(tmp_path / "tiny.png")   # smaller than a crop: skipped
    (tmp_path / "notes.txt").write_text("not an image")
    cfg = {"batch_size": 4, "input_shape": [32, 32, 3], "no_crops_per_image": 2, "inputs": [{"directory": str(tmp_path)}],
           "additional_noise": [5, 40]}
    ds = bfcnn.dataset_builder(cfg)
    assert ds.batch_size == 4 and ds.input_shape == [32, 32, 3] and ds.testing is None     # dataset.py:27-35
    pipe = ds.training
    assert isinstance(pipe, ImagePipeline) and len(pipe.sources) == 6 and pipe.crops_per_image == 2
    crops = pipe._crops(pipe.sources[0], np.random.default_rng(1))
    assert len(crops) == 2 and all(c.shape == (32, 32, 3) and c.dtype == np.uint8 for c in crops)
    assert pipe._crops(str(tmp_path / "tiny.png"), np.random.default_rng(1)) == []
    img = bfcnn.load_image(str(tmp_path / "im0.png"), expand_dims=True)
    assert img.shape == (1, 40, 50, 3) and img.dtype == np.uint8
    assert bfcnn.load_image(str(tmp_path / "im0.png"), image_size=(16, 24), expand_dims=False).shape == (16, 24, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        next(iter(pipe))                                  # the corruption needs the GPU
    with pytest.raises(RuntimeError, match="no CPU path"):
        ds.prepare_data_fn(np.zeros((1, 8, 8, 3), np.uint8))
    # rank sharding: disjoint image sets
    a = ImagePipeline(cfg, ds.noise_config, rank=0, world=2)
    b = ImagePipeline(cfg, ds.noise_config, rank=1, world=2)
    assert len(a) + len(b) <= (6 * 2) // 4 + 1
    syn = bfcnn.dataset_builder({"batch_size": 2, "input_shape": [16, 16, 3], "inputs": [{"synthetic": {"images": 3, "height": 20, "width": 24, "seed": 1}}]})
    assert len(syn.training.sources) == 3 and syn.training.sources[0].shape == (20, 24, 3)
    with pytest.raises(ValueError, match="non directory"):
        bfcnn.dataset_builder({"batch_size": 2, "input_shape": [16, 16, 3], "inputs": [{"directory": str(tmp_path / "empty")}]})
    assert bfcnn.dataset_builder({"batch_size": 2, "input_shape": [16, 16, 3]}).training is None
    with pytest.raises(ValueError, match="color_mode"):
        bfcnn.dataset_builder({"batch_size": 2, "input_shape": [16, 16, 3], "color_mode": "cmyk"})


def test_oracle_multiscales_and_draw_groups():
    """The checker's restatements of utilities.py:625-685 and of the call-level draws of dataset.py:141-187."""
    from oracle import bfcnn_oracle as O
    from oracle import corrupt_oracle as C
    x = np.arange(2 * 6 * 5 * 3, dtype=np.float32).reshape(2, 6, 5, 3) % 251
    s = O.multiscales(x, 2)
    assert [t.shape for t in s] == [(2, 6, 5, 3), (2, 3, 2, 3), (2, 1, 1, 3)]
    assert s[1][0, 0, 0, 0] == np.rint((x[0, 0, 0, 0] + x[0, 0, 1, 0] + x[0, 1, 0, 0] + x[0, 1, 1, 0]) / 4)
    half = np.array([[[[1.0], [2.0]], [[2.0], [1.0]]]], np.float32).repeat(3, -1)      # mean 1.5 -> rounds to 2 (half to even)
    assert O.multiscales(half, 1)[1].tolist() == [[[[2.0, 2.0, 2.0]]]]
    assert O.multiscales(np.full((1, 2, 2, 3), 300.0, np.float32), 1)[1].max() == 255.0   # clip before round
    u8 = np.random.default_rng(0).integers(0, 256, (6, 8, 8, 3), dtype=np.uint8)
    same = u8.copy(); same[:] = u8[0]
    cfg = C.NoiseConfig(random_left_right=False, random_up_down=False, multiplicative_max=0.0, multiplicative_min=0.0)
    # groups of 3: within a group the on/off switch and sigma agree, the per-pixel noise does not
    _, n3 = C.corrupt(same, 5, 0, C.NoiseConfig(**{**cfg.__dict__, "draw_group": 3}))
    for g in (0, 3):
        p = [C.sample_parameters(g, 5, cfg) for _ in range(1)][0]
        grp = n3[g:g + 3] - same[g:g + 3].astype(np.float32)
        on = [bool(np.abs(d).max() > 0) for d in grp]
        assert on == [p["use_add"]] * 3
        if p["use_add"]:
            assert not np.array_equal(grp[0], grp[1])
            assert np.abs(grp).max() <= 2 * float(p["sigma_add"]) + 0.5
    _, nall = C.corrupt(same, 5, 2, C.NoiseConfig(**{**cfg.__dict__, "draw_group": -1}))
    p = C.sample_parameters(2, 5, cfg)
    assert all(bool(np.abs(nall[i] - same[i]).max() > 0) == p["use_add"] for i in range(6))
