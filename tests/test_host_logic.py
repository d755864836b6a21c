"""Host-side logic of the drop-in class that needs no GPU: argument contract, image conversion,
checkpoint discovery, the job splitter and its final gather (gloo, world_size 2)."""
import os
import sys

import numpy as np
import pytest
from PIL import Image

from manga_ocr_b200 import ocr as O
from manga_ocr_b200.splitter import shard_bounds, shard_sizes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_exports_mangaocr():
    import manga_ocr
    assert manga_ocr.MangaOcr is O.MangaOcr
    import inspect
    sig = inspect.signature(O.MangaOcr.__init__)
    names = list(sig.parameters)
    assert names[1:3] == ["pretrained_model_name_or_path", "force_cpu"]
    assert sig.parameters["pretrained_model_name_or_path"].default == "kha-white/manga-ocr-base"
    assert sig.parameters["force_cpu"].default is False


def test_force_cpu_is_refused():
    with pytest.raises(RuntimeError, match="no CPU path"):
        O.MangaOcr(force_cpu=True)


def test_missing_checkpoint_raises(tmp_path, monkeypatch):
    monkeypatch.setenv("HF_HOME", str(tmp_path))
    monkeypatch.delenv("MOCR_WEIGHTS", raising=False)
    with pytest.raises(FileNotFoundError):
        O.MangaOcr()      # the app treats any exception as "engine unavailable" (main_window.py:3396-3398)


def test_find_checkpoint(tmp_path, monkeypatch):
    d = tmp_path / "ckpt"
    d.mkdir()
    (d / "model.safetensors").write_bytes(b"x")
    (d / "vocab.txt").write_text("a\n")
    assert O._find_checkpoint(str(d)) == (str(d / "model.safetensors"), str(d / "vocab.txt"))
    monkeypatch.setenv("HF_HOME", str(tmp_path / "hf"))
    snap = tmp_path / "hf" / "hub" / "models--kha-white--manga-ocr-base" / "snapshots" / "abc"
    snap.mkdir(parents=True)
    (snap / "model.safetensors").write_bytes(b"x")
    assert O._find_checkpoint("kha-white/manga-ocr-base") == (str(snap / "model.safetensors"), None)
    # the snapshot refs/main points to wins over the lexicographically first one, and a snapshot that only
    # ships pytorch_model.bin (what from_pretrained would load) is accepted
    main = snap.parent / "zzz"
    main.mkdir()
    (main / "pytorch_model.bin").write_bytes(b"x")
    refs = snap.parent.parent / "refs"
    refs.mkdir()
    (refs / "main").write_text("zzz\n")
    assert O._find_checkpoint("kha-white/manga-ocr-base") == (str(main / "pytorch_model.bin"), None)


def test_torch_bin_checkpoint_is_read_without_torch(tmp_path):
    """pytorch_model.bin (torch.save zip) -> float32 arrays; written here WITH torch, read by the product WITHOUT it."""
    import numpy as np
    torch = pytest.importorskip("torch")
    from manga_ocr_b200 import weights as W
    sd = {"a.weight": torch.randn(5, 7), "half": torch.randn(3).half(), "bf": torch.randn(4, 4).bfloat16(),
          "position_ids": torch.arange(10)[None], "transposed": torch.randn(6, 4).t(), "param": torch.nn.Parameter(torch.randn(2, 3))}
    torch.save(sd, tmp_path / "pytorch_model.bin")
    out = W.load_torch_bin(str(tmp_path / "pytorch_model.bin"))
    assert "position_ids" not in out
    for k, v in sd.items():
        if v.is_floating_point():
            assert out[k].dtype == np.float32 and np.array_equal(out[k], v.detach().float().numpy()), k
    import pickle

    class Evil:
        def __reduce__(self):
            return (print, ("pwned",))
    import zipfile
    with zipfile.ZipFile(tmp_path / "evil.bin", "w") as z:
        z.writestr("archive/data.pkl", pickle.dumps({"x": Evil()}))
    with pytest.raises(pickle.UnpicklingError):
        W.load_torch_bin(str(tmp_path / "evil.bin"))


def test_image_to_array_modes():
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (9, 7, 3), dtype=np.uint8)
    assert np.array_equal(O.image_to_array(Image.fromarray(rgb)), rgb)
    assert O.image_to_array(Image.fromarray(rgb[..., 0])).shape == (9, 7)
    assert O.image_to_array(Image.fromarray(rgb).convert("RGBA")).shape == (9, 7, 4)
    p = Image.fromarray(rgb).convert("P")
    a = O.image_to_array(p)
    assert a.shape == (9, 7, 3)
    # the engine's luma of the expanded RGB equals Pillow's own convert("L") of the palette image
    from oracle.preprocess_np import rgb_to_l
    assert np.array_equal(rgb_to_l(a), np.asarray(p.convert("L")))
    one = Image.fromarray(rgb[..., 0] > 128)
    assert np.array_equal(rgb_to_l(O.image_to_array(one)), np.asarray(one.convert("L")))


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 64, 512, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _gather_worker(rank, world, port, n_total, T, q):
    import torch.distributed as dist
    from manga_ocr_b200.splitter import gather_ids, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = np.arange(n_total * T, dtype=np.int32).reshape(n_total, T)
    lo, hi = shard_bounds(n_total, world, rank)
    out = gather_ids(full[lo:hi], n_total)
    if rank == 0:
        q.put(bool(np.array_equal(out, full)))
    else:
        q.put(out is None)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 8])
def test_gather_ids_gloo_world2(n_total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_total) % 2000
    ps = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_total, 6, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(res)


def test_generation_config_is_read_from_the_checkpoint_directory(tmp_path):
    """generate() uses the checkpoint's generation settings; no file -> greedy."""
    import json
    from manga_ocr_b200.ocr import GREEDY, _generation_config
    assert _generation_config(None) == GREEDY
    assert _generation_config(str(tmp_path / "model.safetensors")) == GREEDY
    (tmp_path / "config.json").write_text(json.dumps({"num_beams": 2, "length_penalty": 1.5, "model_type": "vision-encoder-decoder"}))
    assert _generation_config(str(tmp_path))["num_beams"] == 2
    # generation_config.json is used wholesale when it exists (not merged over config.json: length_penalty stays at its default)
    (tmp_path / "generation_config.json").write_text(json.dumps({"num_beams": 4, "no_repeat_ngram_size": 3, "early_stopping": True}))
    g = _generation_config(str(tmp_path / "model.safetensors"))
    assert g == {"num_beams": 4, "no_repeat_ngram_size": 3, "length_penalty": 1.0, "early_stopping": True}
