"""Write the three model directories the reference README names
(resnet_color_1x{6,12,18}_bn_16x3x3_256x256_l1_relu; reference README.md:62-64,
setup.py:59-73) with DETERMINISTIC SYNTHETIC weights, in the reference's on-disk
format (SavedModel variables bundle + pipeline.json).  The real pretrained blobs are
absent from the reference snapshot; drop real `saved_model/variables/`
files over these and `load_model(name)` picks them up unchanged.

The repository now ships weights TRAINED with its own training path (tools/train_pretrained.py); this tool leaves a
directory whose pipeline.json says so alone unless called with --force (parity tests and the benchmark take their
deterministic weights from `synthetic_variables` / `synthetic_model`, not from these directories)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from blind_image_denoising_b200.arch import Arch, default_pipeline_config  # noqa: E402
from blind_image_denoising_b200.tensorbundle import write_model_variables  # noqa: E402
from blind_image_denoising_b200.weights import synthetic_variables  # noqa: E402

root = Path(__file__).resolve().parents[1] / "blind_image_denoising_b200"
for n in (6, 12, 18):
    name = f"resnet_color_1x{n}_bn_16x3x3_256x256_l1_relu"
    arch = Arch(no_layers=n)
    d = root / "pretrained" / name
    if (d / "pipeline.json").exists() and "--force" not in sys.argv:
        if str(json.loads((d / "pipeline.json").read_text()).get("weights", "")).startswith("TRAINED"):
            print("kept (trained weights)", d)
            continue
    (d / "saved_model" / "variables").mkdir(parents=True, exist_ok=True)
    write_model_variables(str(d / "saved_model" / "variables"), synthetic_variables(arch, seed=0))
    cfg = default_pipeline_config(arch, name)
    cfg["weights"] = "SYNTHETIC (seed 0): the reference snapshot ships no resnet weights"
    (d / "pipeline.json").write_text(json.dumps(cfg, indent=2))
    (root / "configs" / f"{name}.json").write_text(json.dumps(default_pipeline_config(arch, name), indent=2))
    print("wrote", d)
