"""N3: position-embedding interpolation for 448 px checkpoints, against the reference script's
own arithmetic (mmgpt/utils/interpolate_model.py:11-28, restated inline because the script itself stops in
ipdb and rewrites files)."""
import torch
import torch.nn.functional as F

from merlin_b200.mmgpt.utils.interpolate_model import POS_KEY, interpolate, interpolate_position_embedding


def _reference_lines(old_pos_embed):
    # interpolate_model.py:11-28, verbatim arithmetic
    old_grid_size = round((old_pos_embed.shape[0] - 1) ** 0.5)
    grid_size = old_grid_size * 2
    extra_tokens = 1
    pos_emb_tok, pos_emb_img = old_pos_embed[:extra_tokens], old_pos_embed[extra_tokens:]
    pos_emb_img = pos_emb_img.reshape(1, old_grid_size, old_grid_size, -1).permute(0, 3, 1, 2)
    pos_emb_img = F.interpolate(pos_emb_img, size=(grid_size, grid_size), mode='bicubic', align_corners=True)
    pos_emb_img = pos_emb_img.permute(0, 2, 3, 1).reshape(1, grid_size * grid_size, -1)[0]
    return torch.cat([pos_emb_tok, pos_emb_img], dim=0)


def test_matches_reference_arithmetic_224_to_448():
    torch.manual_seed(0)
    old = torch.randn(1 + 16 * 16, 64)          # CLIP-L/14 at 224 px has a 16 x 16 grid
    new = interpolate_position_embedding(old, 32)
    assert new.shape == (1025, 64)               # the 448 px table of the published weights
    assert torch.equal(new, _reference_lines(old))
    assert torch.equal(new[0], old[0])           # class token untouched
    # align_corners=True keeps the four corner positions exactly
    g_old, g_new = old[1:].view(16, 16, 64), new[1:].view(32, 32, 64)
    for (a, b) in (((0, 0), (0, 0)), ((0, 15), (0, 31)), ((15, 0), (31, 0)), ((15, 15), (31, 31))):
        assert torch.allclose(g_old[a], g_new[b], atol=1e-6)


def test_identity_and_arbitrary_grid():
    old = torch.randn(1 + 24 * 24, 32)           # 336 px
    assert torch.equal(interpolate_position_embedding(old, 24), old)
    assert interpolate_position_embedding(old.to(torch.bfloat16), 32).dtype == torch.bfloat16
    assert interpolate_position_embedding(old, 32).shape == (1025, 32)


def test_file_form_rewrites_the_shard(tmp_path):
    old = torch.randn(1 + 8 * 8, 16)
    torch.save({POS_KEY: old, "other": torch.ones(2)}, tmp_path / "pytorch_model-00002-of-00002.bin")
    interpolate(str(tmp_path))
    sd = torch.load(tmp_path / "pytorch_model-00002-of-00002.bin")
    assert sd[POS_KEY].shape == (1 + 16 * 16, 16) and torch.equal(sd[POS_KEY], _reference_lines(old))
    assert sd[POS_KEY.replace("position_embedding.weight", "position_ids")].shape == (1, 257)
