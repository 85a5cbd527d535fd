"""The zero-edit drop-in recipe of INTEGRATION.md §1, executed: the reference's own scripts/sample_diffusion.py is imported
UNMODIFIED under `PYTHONPATH=frido_b200/compat:<repo>:<reference>` and its import lines (sample_diffusion.py:16-21), its
`load_model_from_config` (:452-457) and its output formatting (`custom_to_np` :115-121) resolve to / work on the
B200-native classes.  CPU only: the script needs /root/reference (build container); `.cuda()` is made an identity, so only
construction + checkpoint loading run here — the sampling flow itself is covered on the GPU by tests/test_gpu_model.py.

Environment stubs (test infrastructure, oracle/shims): pytorch_lightning / omegaconf / kornia are not installed in this
image, and the reference's taming/data/utils.py imports `torch._six` (removed in torch 2; the reference pins torch 1.7) —
a two-line stand-in is injected for it.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FRIDO_REFERENCE", "/root/reference")
COMPAT = os.path.join(ROOT, "frido_b200", "compat")
SHIMS = os.path.join(ROOT, "oracle", "shims")

_IMPORT_LINES = ("from frido.util import log_txt_as_img, exists, default, ismap, isimage, mean_flat, count_params\n"
                 "from frido.models.diffusion.ddim import DDIMSampler\n"
                 "from frido.models.diffusion.plms import PLMSSampler\n"
                 "from frido.util import instantiate_from_config_main as instantiate_from_config\n")


def _run(code, paths):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(paths))
    r = subprocess.run([sys.executable, "-c", code], cwd="/tmp", env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


def test_script_import_lines_without_a_reference_checkout():
    """The exact import lines of sample_diffusion.py:16-19 with only the shims on the path (native helper stand-ins)."""
    code = _IMPORT_LINES + (
        "import torch\n"
        "assert DDIMSampler.__module__ == 'frido_b200.samplers' and PLMSSampler.__module__ == 'frido_b200.samplers'\n"
        "assert exists(0) and not exists(None) and default(None, 3) == 3 and default(None, lambda: 4) == 4\n"
        "x = torch.zeros(2, 3, 4, 4)\n"
        "assert isimage(x) and not ismap(x) and ismap(torch.zeros(2, 5, 4, 4)) and mean_flat(x + 1).tolist() == [1.0, 1.0]\n"
        "assert count_params(torch.nn.Linear(3, 2)) == 8\n"
        "im = log_txt_as_img((64, 32), ['a caption', ['tok', 'ens']])\n"
        "assert tuple(im.shape) == (2, 3, 32, 64) and float(im.max()) <= 1.0 and float(im.min()) >= -1.0\n")
    _run(code, [COMPAT, ROOT])


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scripts")), reason="reference checkout not present")
def test_unmodified_script_loads_a_native_model():
    code = f"""
import importlib.util, sys, types, copy
import torch
six = types.ModuleType('torch._six'); six.string_classes = (str, bytes); sys.modules['torch._six'] = six
torch.nn.Module.cuda = lambda self, *a, **k: self          # no GPU in this container
spec = importlib.util.spec_from_file_location('ref_sample_diffusion', {os.path.join(REF, 'scripts', 'sample_diffusion.py')!r})
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)      # runs the script's own import lines
import frido_b200 as fb
assert m.DDIMSampler is fb.DDIMSampler and m.PLMSSampler is fb.PLMSSampler
# everything else in frido.util is the reference's own code, re-exported
assert m.log_txt_as_img.__module__ == 'frido._reference_util', m.log_txt_as_img.__module__
assert m.custom_collate.__module__ == 'taming.data.utils'
from oracle import synth
g = torch.load({os.path.join(ROOT, 'tests', 'golden', 'tiny2.pt')!r}, weights_only=False)
cfg = copy.deepcopy(g['cfg'])
cfg['params']['cond_stage_config'] = '__is_unconditional__'
cfg['params']['first_stage_config']['params']['ckpt_path'] = None
assert cfg['target'] == 'frido.models.diffusion.frido.FridoDiffusion'
sd = synth.synth_state_dict(g['manifest'], g['seed'])
model = m.load_model_from_config(cfg, sd)                  # sample_diffusion.py:452-457, unmodified
assert type(model) is fb.FridoDiffusion and not model.training
assert type(model.model.diffusion_model) is fb.PyUNetModel and type(model.first_stage_model) is fb.VQModelInterface
k = 'model.diffusion_model.time_embed.0.weight'
assert torch.equal(model.state_dict()[k], sd[k])
# attributes the script reads off the model (sample_diffusion.py:168,183-185,222)
dm = model.model.diffusion_model
assert (dm.num_stage, dm.in_channels, dm.image_size) == (2, 6, 8) and model.cond_stage_key and model.first_stage_key
x = torch.linspace(-1.2, 1.2, 2 * 3 * 4 * 4).view(2, 3, 4, 4)
ref_u8 = m.custom_to_np(x)
assert ref_u8.dtype == torch.uint8 and tuple(ref_u8.shape) == (2, 4, 4, 3)
print('ok')
"""
    out = _run(code, [COMPAT, ROOT, SHIMS, REF])
    assert out.strip().endswith("ok")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scripts")), reason="reference checkout not present")
def test_sample_writer_matches_the_scripts_on_wire_outputs(tmp_path):
    """8f.4: SampleWriter's .npz / PNG files against the unmodified script's own `custom_to_np` + `np.savez` naming
    (sample_diffusion.py:293-301) and `save_logs` (:306-334) on the same fp32 images."""
    import numpy as np
    import torch
    from PIL import Image

    import frido_b200 as fb

    g = torch.Generator().manual_seed(4)
    batches = [torch.randn(3, 3, 16, 16, generator=g) * 0.8 for _ in range(3)]
    names = [[f"img_{b}_{i}.jpg" for i in range(3)] for b in range(3)]
    torch.save(dict(batches=batches, names=names), tmp_path / "in.pt")
    ref_dir = tmp_path / "ref"
    code = f"""
import importlib.util, sys, types, os
import numpy as np, torch
six = types.ModuleType('torch._six'); six.string_classes = (str, bytes); sys.modules['torch._six'] = six
spec = importlib.util.spec_from_file_location('ref_sample_diffusion', {os.path.join(REF, 'scripts', 'sample_diffusion.py')!r})
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
d = torch.load({str(tmp_path / 'in.pt')!r})
os.makedirs({str(ref_dir / 'numpy')!r}, exist_ok=True)
all_images, n_saved = [], 0
for x, fn in zip(d['batches'], d['names']):
    logs = dict(sample=x, file_name=fn)
    n_saved = m.save_logs(logs, {str(ref_dir / 'img')!r}, n_saved=n_saved, keys=['sample'])
    all_images.extend([m.custom_to_np(x)])
all_img = np.concatenate(all_images, axis=0)[:8]
np.savez(os.path.join({str(ref_dir / 'numpy')!r}, 'x'.join(str(v) for v in all_img.shape) + '-samples.npz'), all_img)
n2 = m.save_logs(dict(sample=d['batches'][0]), {str(ref_dir / 'img2')!r}, n_saved=5, keys=['sample'])
assert n2 == 8
"""
    _run(code, [COMPAT, ROOT, SHIMS, REF])

    def u8(x, mode):  # host restatement of frido_to_uint8 (the device kernel is pinned to it in tests/test_gpu_kernels.py)
        if mode == "np":
            return ((x + 1) * 127.5).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
        return torch.from_numpy((255 * ((torch.clamp(x, -1.0, 1.0) + 1.0) / 2.0).permute(0, 2, 3, 1).numpy()).astype(np.uint8))

    w = fb.SampleWriter(logdir=str(tmp_path / "ours" / "img"), nplog=str(tmp_path / "ours" / "numpy"), n_samples=8)
    for x, fn in zip(batches, names):
        w.add_batch(u8(x, "np"), u8(x, "pil"), file_names=fn)
    path = w.finish()
    ref_npz = sorted(os.listdir(ref_dir / "numpy"))
    assert [os.path.basename(path)] == ref_npz == ["8x16x16x3-samples.npz"]
    assert np.array_equal(np.load(path)["arr_0"], np.load(ref_dir / "numpy" / ref_npz[0])["arr_0"])
    ours = sorted(os.listdir(tmp_path / "ours" / "img" / "sample"))
    assert ours == sorted(os.listdir(ref_dir / "img" / "sample")) and len(ours) == 9
    for f in ours:
        a = np.asarray(Image.open(tmp_path / "ours" / "img" / "sample" / f))
        b = np.asarray(Image.open(ref_dir / "img" / "sample" / f))
        assert np.array_equal(a, b), f
    # un-named samples are numbered from the running count (sample_diffusion.py:322-323)
    w2 = fb.SampleWriter(logdir=str(tmp_path / "ours" / "img2"))
    w2.n_saved = 5
    w2.add_batch(u8(batches[0], "np"), u8(batches[0], "pil"))
    w2.finish()
    assert sorted(os.listdir(tmp_path / "ours" / "img2" / "sample")) == sorted(os.listdir(ref_dir / "img2" / "sample"))
