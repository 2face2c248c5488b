"""N2: the data contract either side of the hot path, on CPU — the image-path oracle pinned
against Pillow / the CLIP processor, the collator and the interpair token grammar against the reference's own
classes (imported from /root/reference with its unavailable dependencies stubbed) when the tree is mounted, and
against committed expectations always."""
import importlib.util
import io
import json
import os
import sys
import tarfile
import types

import numpy as np
import pytest
import torch

from oracle import image_oracle as IO

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


@pytest.mark.parametrize("hw", [(480, 640), (336, 336), (100, 57), (1080, 1920), (336, 500), (37, 336), (700, 336)])
def test_resize_oracle_is_bit_exact_with_pillow(hw):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(hw[0] * 7 + hw[1])
    img = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
    img[: hw[0] // 3] = 255                      # saturated regions exercise the clamp
    img[-(hw[0] // 4):, : hw[1] // 2] = 0
    ref = np.asarray(Image.fromarray(img).resize((336, 336)))
    got = IO.resize_bicubic_u8(img, 336)
    assert got.dtype == np.uint8 and got.shape == (336, 336, 3)
    assert np.array_equal(got, ref)


def test_rescale_normalize_matches_the_clip_processor():
    Image = pytest.importorskip("PIL.Image")
    transformers = pytest.importorskip("transformers")
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(336, 336, 3), dtype=np.uint8)
    proc = transformers.CLIPImageProcessor(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
    ref = proc.preprocess(Image.fromarray(img), return_tensors="pt", do_resize=False, do_center_crop=False)["pixel_values"][0]
    got = torch.from_numpy(IO.rescale_normalize(img, proc.image_mean, proc.image_std))
    assert got.shape == ref.shape == (3, 336, 336)
    assert (got - ref).abs().max().item() <= 2e-6
    assert tuple(round(m, 8) for m in proc.image_mean) == CLIP_MEAN and tuple(round(s, 8) for s in proc.image_std) == CLIP_STD


# ----------------------------------------------------------------------------------------- token grammar / collator
class _Tok:
    """word-level tokenizer with the special image tokens, a BOS (id 1) and an EOS, HF call signature"""
    bos_token_id, eos_token, pad_token_id, model_max_length = 1, "</s>", 0, 64

    def __init__(self):
        self.vocab = {"<unk>": 0, "<s>": 1, "</s>": 2, "<im_patch>": 32000, "<im_start>": 32001, "<im_end>": 32002, "\n": 13}

    def convert_tokens_to_ids(self, toks):
        return [self.vocab[t] for t in toks]

    def _split(self, text):
        out, cur = [], ""
        i = 0
        specials = ["<im_patch>", "<im_start>", "<im_end>", "</s>", "\n"]
        while i < len(text):
            for sp in specials:
                if text.startswith(sp, i):
                    if cur.strip():
                        out.extend(cur.split())
                    cur = ""
                    out.append(sp)
                    i += len(sp)
                    break
            else:
                cur += text[i]
                i += 1
        if cur.strip():
            out.extend(cur.split())
        return out

    def __call__(self, text, padding=None, max_length=None, truncation=False):
        ids = [1] + [self.vocab.setdefault(w, 100 + len(self.vocab)) for w in self._split(text)]
        if truncation and max_length is not None:
            ids = ids[:max_length]
        return types.SimpleNamespace(input_ids=ids)


def _reference_classes():
    """InterPairWebDataset / DataCollatorForSupervisedDataset of the real reference, with the modules this
    container lacks (webdataset, megfile, ipdb, loguru, boto3, smart_open, joblib) stubbed."""
    root = "/root/reference"
    if not os.path.isdir(os.path.join(root, "mmgpt")):
        return None
    for name in ("webdataset", "megfile", "ipdb", "loguru", "boto3", "smart_open", "joblib"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.logger = types.SimpleNamespace(info=lambda *a, **k: None, exception=lambda *a, **k: None)
            m.s3_path = types.SimpleNamespace()
            m.Parallel = m.delayed = None
            sys.modules[name] = m
    pkgs = {}
    for pkg in ("mmgpt", "mmgpt.utils", "mmgpt.data", "mmgpt.data.dataset"):
        mod = types.ModuleType(pkg)
        mod.__path__ = []
        pkgs[pkg] = mod
    saved = {k: sys.modules.get(k) for k in list(pkgs) + ["mmgpt.utils.constants", "mmgpt.utils.mm_utils",
                                                           "mmgpt.utils.dist_utils", "mmgpt.data.dataset.base_dataset"]}
    try:
        sys.modules.update(pkgs)
        consts = types.ModuleType("mmgpt.utils.constants")
        src = open(os.path.join(root, "mmgpt/utils/constants.py")).read()
        head = src[: src.index("ROOT_PATH")]            # the token constants; the dataset tables below have the syntax error
        exec(head + "\nPAIR_WEBDATA = {}\n", consts.__dict__)
        sys.modules["mmgpt.utils.constants"] = consts
        dist = types.ModuleType("mmgpt.utils.dist_utils")
        dist.get_world_size = lambda: 1
        sys.modules["mmgpt.utils.dist_utils"] = dist

        def load(name, rel):
            spec = importlib.util.spec_from_file_location(name, os.path.join(root, rel))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            return mod

        load("mmgpt.utils.mm_utils", "mmgpt/utils/mm_utils.py")
        load("mmgpt.data.dataset.base_dataset", "mmgpt/data/dataset/base_dataset.py")
        ds = load("mmgpt.data.dataset.interpair_webdataset", "mmgpt/data/dataset/interpair_webdataset.py")
        col = load("mmgpt.data.collator", "mmgpt/data/collator.py")
        return ds.InterPairWebDataset, col.DataCollatorForSupervisedDataset
    except Exception:
        return None
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _ours(tok, image_token_len=4, use_im_start_end=True):
    from merlin_b200.mmgpt.data.dataset.interpair_webdataset import InterPairWebDataset
    ds = InterPairWebDataset.__new__(InterPairWebDataset)
    ds.tokenizer, ds.image_size, ds.use_im_start_end = tok, 8, use_im_start_end
    ds.im_patch_token, ds.im_start_token, ds.im_end_token = 32000, 32001, 32002
    ds.replace_token = "<im_patch>" * image_token_len
    if use_im_start_end:
        ds.replace_token = "<im_start>" + ds.replace_token + "<im_end>"
    return ds


CASES = [
    [("Track the car <image>", "it moves left <image> then right"), ("Describe <image>\n", "a red car")],
    [("What is this?", "a dog"), ("<image>\nBoxes?", "[001, 002, 003, 004]")],
    [("<image>" + " w" * 30, "x " * 20), ("<image> again", "y " * 30)],          # second pair overflows max length
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_token_processor_matches_the_reference(case):
    tok = _Tok()
    text_list = CASES[case]
    images = [torch.full((3, 8, 8), float(i)) for i in range(4)]
    got = _ours(tok).token_processor(list(text_list), list(images))
    ids, labels = got["input_ids"], got["labels"]
    assert ids.shape == labels.shape and ids.shape[0] <= tok.model_max_length
    # every image token and every prompt token is masked; text tokens are their own targets
    assert bool((labels[(ids >= 32000)] == -100).all())
    assert bool(((labels == ids) | (labels == -100)).all())
    ref_classes = _reference_classes()
    if ref_classes is not None:
        RefDS, _ = ref_classes
        ref = RefDS.__new__(RefDS)
        ours = _ours(tok)
        for k in ("tokenizer", "image_size", "use_im_start_end", "im_patch_token", "im_start_token", "im_end_token",
                  "replace_token"):
            setattr(ref, k, getattr(ours, k))
        want = RefDS.token_processor(ref, list(text_list), list(images))
        assert torch.equal(ids, want["input_ids"]) and torch.equal(labels, want["labels"])
        assert len(got["image"]) == len(want["image"])
        assert all(torch.equal(a, b) for a, b in zip(got["image"], want["image"]))
    if case == 2:
        assert len(got["image"]) == 1            # the images of the dropped pair are clipped


def test_collator_matches_the_reference():
    from merlin_b200.mmgpt.data.collator import DataCollatorForSupervisedDataset
    tok = _Tok()
    tok.model_max_length = 6
    inst = [dict(input_ids=torch.tensor([1, 5, 6]), labels=torch.tensor([-100, 5, 6]), image=[torch.zeros(3, 4, 4)] * 2),
            dict(input_ids=torch.arange(1, 9), labels=torch.arange(1, 9), image=[torch.ones(3, 4, 4)])]
    got = DataCollatorForSupervisedDataset(tokenizer=tok)(inst)
    assert got["input_ids"].shape == (2, 6) and got["labels"][0, 3:].tolist() == [-100] * 3
    assert got["attention_mask"].tolist() == [[True] * 3 + [False] * 3, [True] * 6]
    assert [tuple(i.shape) for i in got["images"]] == [(2, 3, 4, 4), (1, 3, 4, 4)]
    ref_classes = _reference_classes()
    if ref_classes is not None:
        want = ref_classes[1](tokenizer=tok)(inst)
        for k in ("input_ids", "labels", "attention_mask"):
            assert torch.equal(got[k], want[k])
        assert all(torch.equal(a, b) for a, b in zip(got["images"], want["images"]))


def test_local_tar_shards_feed_the_dataset_and_the_collator(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    from merlin_b200.mmgpt.data.collator import DataCollatorForSupervisedDataset
    from merlin_b200.mmgpt.data.dataset.interpair_webdataset import InterPairWebDataset
    shard = tmp_path / "shard-000.tar"
    rng = np.random.default_rng(1)
    with tarfile.open(shard, "w") as tf:
        for r in range(3):
            names = [f"vid{r}-f{j}" for j in range(2)]
            meta = dict(prompt="Track <image> then <image>", txt=f"object {r} moves", image_name_list=names)

            def add(name, data):
                ti = tarfile.TarInfo(name)
                ti.size = len(data)
                tf.addfile(ti, io.BytesIO(data))

            add(f"rec{r}.json", json.dumps(meta).encode())
            for j in range(2):
                buf = io.BytesIO()
                Image.fromarray(rng.integers(0, 256, (20, 30, 3), dtype=np.uint8)).save(buf, format="JPEG")
                add(f"rec{r}.f{j}.jpg", buf.getvalue())
    tok = _Tok()
    tok.model_max_length = 256
    cfg = dict(image_token_len=4, use_im_start_end=True, image_size=8, image_processor=None)
    loader = lambda data: torch.from_numpy(IO.image_preprocess(                      # the oracle stands in for the GPU kernel
        np.asarray(Image.open(io.BytesIO(data)).convert("RGB")), 8, CLIP_MEAN, CLIP_STD))
    ds = InterPairWebDataset([str(shard)], tok, cfg, merge_round=2, image_loader=loader)
    assert len(ds) == 1
    sample = ds[0]
    assert len(sample["image"]) == 4 and sample["image"][0].shape == (3, 8, 8)
    assert int((sample["input_ids"] == 32000).sum()) == 4 * 4                         # 4 images x image_token_len
    batch = DataCollatorForSupervisedDataset(tokenizer=tok)([sample, ds[1]])
    assert batch["input_ids"].shape[0] == 2 and batch["images"][0].shape == (4, 3, 8, 8)


def test_build_dataloader_contract(tmp_path):
    """mmgpt/data/builder.py:12-81 mirror: same return keys; local shard globs with an optional @merge_round;
    dataset kinds / image modes outside the mirrored slice fail loudly."""
    from merlin_b200.mmgpt.data.builder import build_dataloader
    from merlin_b200.mmgpt.data.collator import DataCollatorForSupervisedDataset
    shard = tmp_path / "a-000.tar"
    with tarfile.open(shard, "w") as tf:
        for r in range(5):
            data = json.dumps(dict(prompt="p <image>", txt="t", image_name_list=[f"x-f{r}"])).encode()
            ti = tarfile.TarInfo(f"r{r}.json")
            ti.size = len(data)
            tf.addfile(ti, io.BytesIO(data))
    tok = _Tok()
    base = dict(image_token_len=4, use_im_start_end=True, image_size=8, image_processor=None, image_aspect_ratio="resize")
    out = build_dataloader(tok, types.SimpleNamespace(interpair_webdatasets=str(tmp_path / "a-*.tar") + "@2", **base))
    assert set(out) == {"train_dataset", "eval_dataset", "data_collator"} and out["eval_dataset"] is None
    assert isinstance(out["data_collator"], DataCollatorForSupervisedDataset)
    assert len(out["train_dataset"]) == 5 // 2 and out["train_dataset"].merge_round == 2
    two = build_dataloader(tok, types.SimpleNamespace(
        interpair_webdatasets=str(tmp_path / "a-*.tar") + "+" + str(tmp_path / "a-*.tar") + "@5", **base))
    assert len(two["train_dataset"]) == 5 + 1
    with pytest.raises(FileNotFoundError):
        build_dataloader(tok, types.SimpleNamespace(interpair_webdatasets=str(tmp_path / "none-*.tar"), **base))
    with pytest.raises(NotImplementedError):
        build_dataloader(tok, types.SimpleNamespace(interpair_webdatasets=str(shard), conversation_datasets="llava665k", **base))
    with pytest.raises(NotImplementedError):
        build_dataloader(tok, types.SimpleNamespace(interpair_webdatasets=str(shard), **dict(base, image_aspect_ratio="keep")))
    with pytest.raises(ValueError):
        build_dataloader(tok, types.SimpleNamespace(interpair_webdatasets="", **base))


def _write_text_shards(tmp_path, n_shards=8, per_shard=5):
    paths = []
    for s in range(n_shards):
        p = tmp_path / f"s-{s:03d}.tar"
        with tarfile.open(p, "w") as tf:
            for r in range(per_shard):
                data = json.dumps(dict(prompt=None, txt=f"rec {s} {r}", image_name_list=[], uid=s * 100 + r)).encode()
                ti = tarfile.TarInfo(f"r{s}_{r}.json")
                ti.size = len(data)
                tf.addfile(ti, io.BytesIO(data))
        paths.append(str(p))
    return paths


def test_interpair_dataset_shards_by_rank_and_worker_and_shuffles_with_fixed_seeds(tmp_path, monkeypatch):
    """The reference pipeline (interpair_webdataset.py:30-41): InfiniteShardList(seed 3407) -> shuffle(world_size) ->
    split_by_node -> split_by_worker -> tarfile_to_samples -> shuffle(1000, seed 42).  Under data parallelism every
    rank (and every DataLoader worker) must read DIFFERENT shards; same seeds -> same order on every run."""
    from merlin_b200.mmgpt.data.dataset.interpair_webdataset import InterPairWebDataset, shuffle_buffer
    import random
    shards = _write_text_shards(tmp_path)
    cfg = dict(image_token_len=4, use_im_start_end=True, image_size=8, image_processor=None)

    def first_uids(n, **kw):
        ds = InterPairWebDataset(shards, _Tok(), cfg, data_size=40, **kw)
        it = ds.inner_iter
        return [next(it)["json"]["uid"] for _ in range(n)]

    # one epoch = 40 records over 8 shards; with 2 ranks each rank owns 4 shards per epoch
    r0 = first_uids(20, rank=0, world_size=2, shuffle=False)
    r1 = first_uids(20, rank=1, world_size=2, shuffle=False)
    assert not set(r0) & set(r1) and len(set(r0) | set(r1)) == 40          # disjoint, together the whole epoch
    assert {u // 100 for u in r0} == {0, 2, 4, 6} and {u // 100 for u in r1} == {1, 3, 5, 7}
    # torchrun environment is picked up when rank / world_size are not given
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    assert first_uids(20, shuffle=False) == r1
    monkeypatch.delenv("RANK")
    monkeypatch.delenv("WORLD_SIZE")
    # DataLoader workers split the rank's shards again (split_by_worker)
    import torch.utils.data as tud
    seen = {}
    for wid in range(2):
        monkeypatch.setattr(tud, "get_worker_info", lambda wid=wid: types.SimpleNamespace(id=wid, num_workers=2))
        seen[wid] = first_uids(10, rank=0, world_size=2, shuffle=False)
    monkeypatch.undo()
    assert not set(seen[0]) & set(seen[1]) and set(seen[0]) | set(seen[1]) == set(r0)
    # shuffled: deterministic across runs, a permutation that differs from file order, ranks still disjoint per epoch
    a, b = first_uids(60, rank=0, world_size=1), first_uids(60, rank=0, world_size=1)
    assert a == b and a[:40] != sorted(a[:40])
    s0 = {u // 100 for u in first_uids(2000, rank=0, world_size=2)}
    assert len(s0) == 8                                                     # over several epochs a rank sees every shard
    # the shuffle filter itself: a permutation of a finite stream, bounded look-ahead
    out = list(shuffle_buffer(iter(range(50)), 8, random.Random(42)))
    assert sorted(out) == list(range(50)) and out != list(range(50))
    assert all(x <= i + 8 for i, x in enumerate(out))
