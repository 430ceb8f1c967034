"""GPU parity at the BASELINE.json config sizes and on the reference's REAL data (round-2 additions).

Every comparison is CUDA output (through the C ABI) against the CPU oracle (oracle/), never against another kernel:
  * configs[1] shape: 65 536-tile dictionary x 16 palettes, a 512-row sample of a 432 000-tile batch through oracle.match_tiles
  * configs[2] slice: one tensor-core assignment against 262 144 centroids, 2 000 points against an exact f64 scan
  * configs[4] shape: 32 palettes x 256 colours, 2 048 (tile, palette) pairs, both ditherers
  * configs[0]: 320x180x24 end to end (the 180 rows zero-padded to 184 inside encode()) against oracle/pipeline.py
  * the dithered tiles and palettes of docs/demo/city_cif.gtm and football_cif.gtm (tests/golden/demo_tiles.npz)
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import rand_tiles, ROOT
from tiler_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _u32(a):
    return np.asarray(a).view(np.uint32) if np.asarray(a).dtype == np.int32 else np.asarray(a)


# ------------------------------------------------------------------ configs[1]: 65 536 x 16 x 16, k = 64, extended re-rank
def test_config_b_match_sample_against_oracle(tm, oracle):
    from tiler_b200.encoder import TilingEncoder
    clip = synth.make_clip(1280, 720, 6, cut_every=0, seed=synth.SEED + 5)
    tiles = synth.clip_to_tiles(clip).reshape(-1, 64)                          # 86 400 source tiles
    enc = TilingEncoder(palette_size=16, palette_count=16)
    canon, flags = enc.load_tiles(tiles[None])
    enc.reduce_sample(canon, flags, 65536)
    enc.prepare_palettes()
    enc.dither()
    enc.prepare_reconstruct()
    didx, dpal, pal = np.asarray(enc.tile_idx), np.asarray(enc.tile_pal), np.asarray(enc.palettes)
    assert didx.shape == (65536, 64) and pal.shape == (16, 16)
    dict_feat = oracle.features_from_pal(didx, dpal, pal)
    assert np.array_equal(enc.matcher.dict_features(), dict_feat)              # PrepareReconstruct at full dictionary size
    # a full keyframe-sequence batch (432 000 tiles = 5 frames repeated): every k-NN wave shape of the bench step
    batch = np.ascontiguousarray(np.tile(canon.reshape(-1, 64), (5, 1)))
    assert batch.shape[0] == 432000
    ti, pi, er = enc.matcher.match_rgb(batch, 64)
    sel = np.linspace(0, batch.shape[0] - 1, 512).astype(np.int64)
    qf = oracle.features_from_rgb(batch[sel])
    ot, op, oe = oracle.match_tiles(qf, dict_feat, didx, dpal, pal, k=64, extended=True)
    assert np.array_equal(_u32(er)[sel], oe)
    assert np.array_equal(ti[sel], ot) and np.array_equal(pi[sel], op)
    # the raw 64-NN of the same rows: same distance multiset, same index set wherever the 64th distance is not tied
    knn = tm.KnnShort(dict_feat)
    gi, gd = knn.search(qf, 64)
    oi, od = oracle.knn_short(dict_feat, qf, 64)
    assert np.array_equal(_u32(gd), od) and np.array_equal(gi, oi)
    knn.close()
    enc.finish_reconstruct()


# ------------------------------------------------------------------ configs[2] slice: K = 262 144 centroids
def test_config_c_assignment_slice_against_exact_f64(tm, oracle):
    rng = np.random.default_rng(2024)
    K, n = 262144, 148 * 128 * 2
    centres = synth.random_features(K, 11, adversarial=True).astype(np.float64)
    centres += rng.uniform(-0.5, 0.5, size=centres.shape)                      # f64 centroids, not on the int16 lattice
    pick = rng.integers(0, K, size=n)
    x = np.clip(np.rint(centres[pick] + rng.normal(0, 25, size=(n, 192))), -32768, 32767).astype(np.int16)
    labels = np.full(n, -1, np.int32)
    labels, sums, counts, changed, inertia = tm.kmeans_partial_step_i16(x, centres, labels)
    sel = rng.choice(n, size=2000, replace=False)
    oi, od = oracle.knn_double(centres, x[sel].astype(np.float64))
    assert np.array_equal(np.asarray(labels)[sel], oi)                          # exact f64 decision, first minimum
    assert changed == n and int(np.asarray(counts).sum()) == n
    # per-cluster partial sums of the shard: every point lands in its label's row
    want = np.zeros((K, 192)); np.add.at(want, np.asarray(labels), x.astype(np.float64))
    assert np.array_equal(np.asarray(sums), want)                               # integer-valued sums: exact in any order


# ------------------------------------------------------------------ configs[4] shape: 32 palettes x 256 colours
@pytest.mark.parametrize("use_tk", [True, False])
def test_config_e_dither_256_colours_against_oracle(tm, oracle, use_tk):
    frame = synth.pack_rgb(synth.make_clip(512, 256, 1, cut_every=0, seed=77))[0]
    tiles = synth.frame_to_tiles(frame)                                        # 2 048 tiles of a generator frame
    canon, flags = tm.mirror_canonicalise(tiles)
    bands = (np.arange(len(tiles)) * 32 // len(tiles)).astype(np.int32)        # 32 spatial bands -> 32 palettes
    pal, _ = tm.palquant_kmeans(canon, bands, 32, 256, seed=7)
    for p in range(0, 32, 5):
        want, _ = oracle.quantize_palette(canon[bands == p].reshape(-1), 256, seed=7)
        assert np.array_equal(np.asarray(pal)[p], want)
    rng = np.random.default_rng(9)
    pair_tile = rng.integers(0, len(tiles), size=2048).astype(np.int32)
    pair_pal = rng.integers(0, 32, size=2048).astype(np.int32)
    got = tm.dither(canon, flags, pair_pal, pal, use_thomas_knoll=use_tk, y2_mixed_colors=4, pair_tile=pair_tile)
    want = oracle.dither(canon, flags, pair_pal, np.asarray(pal), use_tk=use_tk, y2_mixed_colors=4, pair_tile=pair_tile)
    assert np.array_equal(got, want)


# ------------------------------------------------------------------ configs[0]: 320x180x24, one keyframe sequence, end to end
def test_config_a_encode_320x180_against_oracle_pipeline(tm, oracle):
    """The whole encode against oracle/pipeline.py (which imports nothing from the product).  Stated tolerance: decoded-frame
    PSNR within 0.05 dB (north star); the stages are deterministic and bit-exact, so the comparison is equality of the
    dictionary, palettes, tilemap and reconstructed frames, and of the frames the stream decodes to."""
    from oracle import pipeline as P
    from tiler_b200 import gtm
    from tiler_b200.encoder import TilingEncoder, psnr_rgb
    w, h, n = 320, 180, 24
    frames = synth.pack_rgb(synth.make_clip(w, h, n, cut_every=0, seed=synth.SEED))
    seqs = [(0, n - 1)]
    # LoadDefaultSettings (:3817-3845): PaletteSize 16, PaletteCount 1024, QualityBasedTileCount 7.0 -> 7 * EqualQualityTileCount(22080) = 15008
    n_pal, pal_size, tile_count = 1024, 16, 15008
    enc = TilingEncoder(palette_size=pal_size, palette_count=n_pal, seed=0x42381337)
    res = enc.encode(frames, seqs, tile_count=tile_count, radius=32)
    ref = P.encode(frames, seqs, tile_count, n_pal, pal_size, 0x42381337, radius=32)
    assert (ref["tw"], ref["th"]) == (40, 23)                                   # (h - 1) div 8 + 1 (:1776)
    assert enc.reduce_threshold == ref["threshold"]
    assert np.array_equal(res["palettes"], ref["palettes"])
    assert np.array_equal(res["tiles"], ref["tiles"]) and np.array_equal(res["use_count"], ref["use_count"])
    tmap = res["tilemap"]
    for key in ("tile_idx", "pal_idx", "pred_x", "pred_y", "is_pred"):
        assert np.array_equal(np.asarray(tmap[key]).reshape(n, -1), np.asarray(ref[key]).reshape(n, -1)), key
    assert np.array_equal(res["recon"], ref["recon"])
    decoded, hdr = gtm.decode_gtm(res["gtm"])
    assert (hdr["width"], hdr["height"], hdr["frame_count"]) == (320, 184, n)
    assert np.array_equal(decoded, ref["recon"])
    src = np.zeros((n, 184, 320), np.int32); src[:, :180] = frames
    assert abs(psnr_rgb(decoded, src) - psnr_rgb(ref["recon"], src)) <= 0.05


# ------------------------------------------------------------------ the reference's real dithered tiles and palettes
@pytest.mark.parametrize("name", ["city_cif", "football_cif"])
def test_demo_stream_tiles_features_knn_and_matcher(tm, oracle, name):
    z = np.load(os.path.join(GOLD, "demo_tiles.npz"))
    didx, dpal, pal = z[name + "_tiles"], z[name + "_tile_pal"], z[name + "_palettes"]
    n = len(didx)
    dict_feat = oracle.features_from_pal(didx, dpal, pal)
    assert np.array_equal(tm.features_from_pal(didx, dpal, pal), dict_feat)     # all 20 000 real tiles, bit-exact
    # queries: the RGB rendering of every 9th tile under ANOTHER palette of the same stream (what the extended re-rank sees)
    rng = np.random.default_rng(4)
    qsel = np.arange(0, n, 9)
    qpal = rng.integers(0, len(pal), size=len(qsel))
    q_rgb = pal[qpal[:, None], didx[qsel]].astype(np.int32)
    qf = oracle.features_from_rgb(q_rgb)
    assert np.array_equal(tm.features_from_rgb(q_rgb), qf)
    knn = tm.KnnShort(dict_feat)
    for k in (1, 64):
        gi, gd = knn.search(qf[:700], k)
        oi, od = oracle.knn_short(dict_feat, qf[:700], k)
        assert np.array_equal(_u32(gd), od), k
        ties = (od[:, -1:] == od).sum(1) > 1 if k > 1 else np.zeros(len(od), bool)
        assert np.array_equal(gi[~ties], oi[~ties]) and np.array_equal(gi, oi), k   # identical tie order as well ((distance, index))
    knn.close()
    m = tm.Matcher(didx, dpal, pal, extended=True)
    ti, pi, er = m.match_rgb(q_rgb[:700], 64)
    ot, op, oe = oracle.match_tiles(qf[:700], dict_feat, didx, dpal, pal, k=64, extended=True)
    assert np.array_equal(_u32(er), oe) and np.array_equal(ti, ot) and np.array_equal(pi, op)
    m.close()
    # re-dithering the rendered real tiles against their own real palette (Thomas Knoll and Yliluoma)
    flags = np.zeros(256, np.uint8)
    rgb_own = pal[dpal[:256, None], didx[:256]].astype(np.int32)
    for use_tk in (True, False):
        assert np.array_equal(tm.dither(rgb_own, flags, dpal[:256], pal, use_thomas_knoll=use_tk),
                              oracle.dither(rgb_own, flags, dpal[:256], pal, use_tk=use_tk))


# ------------------------------------------------------------------ DoPalettization chain and its coreset stand-in
def test_palettization_chain_against_oracle(tm, oracle):
    from oracle import pipeline as P
    from tiler_b200.encoder import TilingEncoder
    tiles = rand_tiles(3000, 31)
    use = np.random.default_rng(2).integers(1, 40, size=3000).astype(np.int32)
    feats = np.stack([oracle.tile_features_f64(t, oracle.PVS_WEIGHTED_SPE_DCT, True) for t in tiles])
    got_c, got_w = tm.coreset_weighted(feats, use.astype(np.float64), 96, seed=0x42381337)
    ref_c, ref_w = oracle.coreset_weighted(feats, use.astype(np.float64), 96, seed=0x42381337)
    assert np.array_equal(got_c, ref_c) and np.array_equal(got_w, ref_w)       # the BICO stand-in, bit for bit
    # drop-in BICO symbols: the same summary through bico_create / insert_line / get_results
    b = tm.Bico(192, 3000, 96, 32, 96, 0x42381337)
    for row, wt in zip(feats, use):
        b.insert_line(row, float(wt))
    bc, bw = b.get_results()
    b.destroy()
    assert np.array_equal(bc, ref_c) and np.array_equal(bw, ref_w)
    for n_pal in (12, 1):
        enc = TilingEncoder(palette_size=16, palette_count=n_pal, seed=0x42381337)
        enc.tiles, enc.tile_flags, enc.use_count = tiles, np.zeros(3000, np.uint8), use
        enc.prepare_palettes()
        # the GPU features are within 1e-5 relative of the oracle's (libm pow), so the chain is compared from the oracle's own
        # features only where no decision sits inside that tolerance: labels must agree on >= 99.9 % of the tiles
        want = P.palettize(tiles, use, n_pal, 0x42381337)
        agree = (np.asarray(enc.tile_pal) == want).mean()
        assert agree >= 0.999, agree
    # coreset smaller than the palette count: identity mapping (:4214-4219)
    enc = TilingEncoder(palette_size=16, palette_count=64, seed=1)
    enc.tiles, enc.tile_flags, enc.use_count = tiles[:40], np.zeros(40, np.uint8), use[:40]
    enc.prepare_palettes()
    assert enc.coreset_size == 40 and len(np.unique(np.asarray(enc.tile_pal))) == 40


# ------------------------------------------------------------------ the CUDA-core motion kernel (QuickTest prune), bit-exact too
def test_motion_scalar_kernel_in_subprocess():
    """abi.cu reads TM_MOTION_SCALAR once per process: run the motion / Reconstruct parity tests again in a child process with
    the CUDA-core kernel (motion_search_kernel: QuickTestEuclideanDCTPtr prune, utils.pas:755-759) selected."""
    env = dict(os.environ, TM_MOTION_SCALAR="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_core.py"), "-m", "gpu", "-q", "-x", "-k",
                        "motion_search_bit_exact or reconstruct_sequence_bit_exact or predict_motion_frame"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


# ------------------------------------------------------------------ one process driving two devices (per-device state)
def test_two_devices_in_one_process(tm, oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process (gpurun --gpus 2)")
    d = synth.random_features(3000, 1)
    q = synth.random_features(500, 2)
    oi, od = oracle.knn_short(d, q, 64)
    tiles = rand_tiles(64, 3)
    want_f = oracle.features_from_rgb(tiles)
    for dev in (0, 1, 0):
        dd, qq = torch.from_numpy(d).to(f"cuda:{dev}"), torch.from_numpy(q).to(f"cuda:{dev}")
        knn = tm.KnnShort(dd)
        gi, gd = knn.search(qq, 64)
        assert np.array_equal(gi.cpu().numpy(), oi) and np.array_equal(_u32(gd.cpu().numpy()), od)
        knn.close()
        f = tm.features_from_rgb(torch.from_numpy(tiles).to(f"cuda:{dev}"))
        assert np.array_equal(f.cpu().numpy(), want_f)
        f64 = tm.features_f64(torch.from_numpy(tiles).to(f"cuda:{dev}"))
        assert f64.device.index == dev


def test_sharded_encode_equals_single_process_encode():
    """Two ranks (torchrun, NCCL): the sharded encode -- frames uploaded n / N per rank and all-gathered, PSNRs / tilemaps / use counts
    exchanged as device tensors, LZMA chunks compressed per rank -- gives the single-process stream byte for byte, in both feature
    modes, with uneven sequences and a frame count that is not a multiple of the world size."""
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "sharded_encode_check.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    for mode in ("fast", "exact"):
        assert res[mode]["equal"] and res[mode]["tilemap_equal"], res


# ------------------------------------------------------------------ fast (separable f64) sliding-window features
@pytest.mark.parametrize("w,h", [(96, 64), (100, 52), (320, 184), (15, 8)])
def test_sliding_features_fast_mode_within_contract(tm, oracle, w, h):
    """tm_set_feature_mode(1): <= 1 LSB per coefficient and <= 1e-3 of the coefficients differing from the bit-exact features
    (SURVEY "Parity contract"); the bit-exact mode is restored afterwards and still equals the oracle."""
    frame = synth.pack_rgb(synth.make_clip(max(w, 16), max(h, 16), 1, cut_every=0, seed=w * 1000 + h, n_sprites=3))[0][:h, :w]
    frame = np.ascontiguousarray(frame)
    want = oracle.sliding_features(frame)
    prev = tm.set_feature_mode(tm.FEATURES_FAST)
    try:
        got = tm.sliding_features(frame)
    finally:
        tm.set_feature_mode(prev)
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1
    assert (d != 0).mean() <= 1e-3, (d != 0).mean()
    assert np.array_equal(tm.sliding_features(frame), want)


@pytest.mark.parametrize("w,h", [(320, 184), (104, 56), (1280, 720)])
def test_fast_mode_fused_limb_rows_equal_the_two_pass_path(tm, oracle, w, h):
    """In the fast feature mode tm_predict_motion_frame lets the sliding-window kernel write the tensor-core search's candidate
    operands (limb rows + norms at padded positions) itself.  The result must equal the two-pass path -- fast int16 features
    (tm_sliding_features) split by the search (tm_motion_search) -- bit for bit: same features, only the intermediate is gone.
    Widths whose padded pitch exceeds 8 * tiles (104 -> pitch 128) exercise the all-padding column blocks."""
    tw, th = w // 8, h // 8
    clip = synth.pack_rgb(synth.make_clip(w, h, 2, cut_every=0, seed=w + h, n_sprites=6))
    prev_frame, cur = np.ascontiguousarray(clip[0]), np.ascontiguousarray(clip[1])
    tiles = np.ascontiguousarray(cur.reshape(th, 8, tw, 8).transpose(0, 2, 1, 3).reshape(-1, 64))
    canon, flags = tm.mirror_canonicalise(tiles)
    mode = tm.set_feature_mode(tm.FEATURES_FAST)
    try:
        x1, y1, e1 = tm.predict_motion_frame(prev_frame, canon, flags, tw, th, 16)
        dcts = tm.sliding_features(prev_frame)
        x2, y2, e2 = tm.motion_search(tm.features_from_rgb_mirrored(canon, flags), tw, th, dcts, 16)
    finally:
        tm.set_feature_mode(mode)
    assert np.array_equal(np.asarray(e1).view(np.uint32), np.asarray(e2).view(np.uint32))
    assert np.array_equal(x1, x2) and np.array_equal(y1, y2)


def test_fast_features_encode_psnr_within_tolerance(tm, oracle):
    """End to end with the fast sliding-window features: decoded-frame PSNR within 0.05 dB of the bit-exact encode; tile indices
    that differ are counted (they sit inside the distance tolerance of +-1 LSB features)."""
    from tiler_b200 import gtm
    from tiler_b200.encoder import TilingEncoder, psnr_rgb
    w, h, n = 160, 96, 8
    frames = synth.pack_rgb(synth.make_clip(w, h, n, cut_every=4, seed=77, n_sprites=5))
    seqs = [(0, 3), (4, 7)]
    exact = TilingEncoder(palette_size=16, palette_count=4, seed=3).encode(frames, seqs, tile_count=400)
    fast = TilingEncoder(palette_size=16, palette_count=4, seed=3, feature_mode="fast").encode(frames, seqs, tile_count=400)
    assert tm.set_feature_mode(tm.FEATURES_EXACT) == tm.FEATURES_EXACT          # encode() restored the process-wide mode
    pe = psnr_rgb(gtm.decode_gtm(exact["gtm"])[0], frames)
    pf = psnr_rgb(gtm.decode_gtm(fast["gtm"])[0], frames)
    assert abs(pe - pf) <= 0.05, (pe, pf)
    differing = (np.asarray(exact["tilemap"]["is_pred"]) != np.asarray(fast["tilemap"]["is_pred"])).mean()
    print(f"fast features: PSNR {pf:.4f} dB vs {pe:.4f} dB, {differing:.2%} of the tilemap items change their predicted flag")
    assert differing <= 0.02


# ------------------------------------------------------------------ persistent shard handle of the sharded Lloyd loop
def test_kmeans_i16_shard_handle_equals_single_call_fit(tm, oracle):
    """tiler_b200.dist.kmeans_fit_i16_sharded (limb rows split once, stopping test read every 4 iterations, nothing copied to the
    host per step) against tm_kmeans_fit_i16 and the oracle's f64 Lloyd from the same initial centroids."""
    import torch
    from tiler_b200 import dist as tdist
    x = synth.random_features(6000, 9, adversarial=True)
    init = x[:96].astype(np.float64)
    ol, oc, oin, oit = oracle.kmeans_lloyd(x.astype(np.float64), init, max_iter=300)
    gl, gc, gin, git, _ = tm.kmeans_fit_i16(x, 96, init, max_iter=300)
    assert np.array_equal(gl, ol) and np.array_equal(gc, oc)
    xs, ini = torch.from_numpy(x).cuda(), torch.from_numpy(init).cuda()
    for every in (1, 4):
        tms = {}
        sl, sc, sin, sit = tdist.kmeans_fit_i16_sharded(xs, ini, max_iter=300, check_every=every, timings=tms)
        assert np.array_equal(sl.cpu().numpy(), ol) and np.array_equal(sc.cpu().numpy(), oc)
        assert abs(sin - oin) <= 1e-9 * oin and oit <= sit < oit + every
        assert len(tms["assign_ms"]) == sit + 1


# ------------------------------------------------------------------ MergeTiles hook: a k-means-built dictionary (north star stage 3)
def test_cluster_dictionary_through_merge_tiles_hook(tm, oracle):
    from oracle import pipeline as P
    from tiler_b200.encoder import TilingEncoder
    frames = synth.pack_rgb(synth.make_clip(128, 96, 4, cut_every=0, seed=19, n_sprites=4))
    _, canon, flags, tw, th = P.load(frames)
    psnr = P.predict_motion(P.pad_frames(frames)[0], canon, flags, tw, th, 32)
    enc = TilingEncoder(palette_size=16, palette_count=2)
    tmap = enc.reduce(canon, flags, psnr, [0], 300)
    d_tiles, d_use = np.asarray(enc.tiles).copy(), np.asarray(enc.use_count).copy()
    n, k = len(d_tiles), 60
    new_map = enc.cluster_dictionary(k, tmap)
    # the same clustering restated with the oracle: f64 Lloyd from the k most used tiles, nearest member kept, MergeTiles
    feats = oracle.features_from_rgb(d_tiles).astype(np.float64)
    lab, cen, _, _ = oracle.kmeans_lloyd(feats, feats[:k].copy(), max_iter=300)
    d = ((feats - cen[lab]) ** 2).sum(1)
    best = np.array([min(np.flatnonzero(lab == c), key=lambda i: (d[i], i)) if (lab == c).any() else -1 for c in range(k)])
    r_tiles, r_use, r_map = P.merge_tiles(d_tiles, d_use, lab, best, tmap)
    assert len(r_tiles) <= k and r_use.sum() == d_use.sum()
    assert np.array_equal(np.asarray(enc.tiles), r_tiles) and np.array_equal(enc.use_count, r_use) and np.array_equal(new_map, r_map)
    # the clustered dictionary still encodes: palettes, dithering and the matcher accept it
    enc.prepare_palettes(); enc.dither(); enc.prepare_reconstruct()
    ti, pi, er = enc.reconstruct(canon[0])
    assert ti.min() >= 0 and ti.max() < len(r_tiles)
    enc.finish_reconstruct()


# ------------------------------------------------------------------ mirror-variant search (north star / configs[4])
@pytest.mark.parametrize("extended", [True, False])
def test_mirror_variant_search_against_oracle(tm, oracle, extended):
    rng = np.random.default_rng(61)
    tiles = rand_tiles(900, 71)
    canon, flags = tm.mirror_canonicalise(tiles)
    pal = np.stack([oracle.quantize_palette(canon[p::4].reshape(-1), 16, seed=3)[0] for p in range(4)])
    dpal = (np.arange(600) % 4).astype(np.int32)
    didx = oracle.dither(canon[:600], flags[:600], dpal, pal)
    # queries: dictionary source tiles deliberately stored in a NON-canonical orientation, plus unrelated tiles
    q = canon[300:900].copy().reshape(-1, 8, 8)
    flip = rng.integers(0, 4, size=len(q))
    for i, v in enumerate(flip):
        if v & 1: q[i] = q[i][:, ::-1]
        if v & 2: q[i] = q[i][::-1, :]
    q = np.ascontiguousarray(q.reshape(-1, 64))
    m = tm.Matcher(didx, dpal, pal, extended=extended)
    ti, pi, er, var = m.match_rgb_mirrors(q)
    m.close()
    dict_feat = oracle.features_from_pal(didx, dpal, pal)
    res = []
    for v in range(4):
        qf = oracle.features_from_rgb_mirrored(q, np.full(len(q), v, np.uint8))
        assert np.array_equal(tm.features_from_rgb_mirrored(q, np.full(len(q), v, np.uint8)), qf)
        res.append(oracle.match_tiles(qf, dict_feat, didx, dpal, pal, k=64, extended=extended))
    e4 = np.stack([r[2] for r in res]).astype(np.int64)
    bv = e4.argmin(0)                                             # first minimum: the canonical orientation keeps ties
    rows = np.arange(len(q))
    assert np.array_equal(var, bv)
    assert np.array_equal(_u32(er), e4[bv, rows].astype(np.uint32))
    assert np.array_equal(ti, np.stack([r[0] for r in res])[bv, rows]) and np.array_equal(pi, np.stack([r[1] for r in res])[bv, rows])
    assert (var[:300][flip[:300] != 0] != 0).mean() > 0.5            # re-flipped dictionary tiles are found through a mirror variant


# ------------------------------------------------------------------ drop-in per-query searches from many host threads
def test_dropin_search_micro_batching_from_threads(tm, oracle):
    """ann_kdtree_short_search / _search_multi and ann_kdtree_search called like the host's MTProcs workers do (one query per
    call, many threads): results equal the oracle's and the calls are combined into far fewer launches than queries."""
    from concurrent.futures import ThreadPoolExecutor
    d = synth.random_features(5000, 21)
    q = synth.random_features(640, 22)
    oi1, od1 = oracle.knn_short(d, q, 1)
    oi64, od64 = oracle.knn_short(d, q, 64)
    t = tm.AnnKdTreeShort(d)

    def work(i):
        idx, err = t.search(q[i])
        mi, me = t.search_multi(q[i], 64)
        return idx, err, mi, me
    with ThreadPoolExecutor(max_workers=32) as pool:
        res = list(pool.map(work, range(len(q))))
    assert [r[0] for r in res] == list(oi1[:, 0]) and [r[1] for r in res] == list(od1[:, 0])
    assert all(np.array_equal(r[2], oi64[i]) and np.array_equal(r[3], od64[i]) for i, r in enumerate(res))
    nq, nb = t.rendezvous_stats()
    print(f"rendezvous: {nq} queries in {nb} batched launches ({nq / nb:.1f} per launch)")
    assert nq == 2 * len(q) and nb < nq
    t.destroy()
    pts = np.random.default_rng(5).normal(size=(300, 24))
    qs = np.random.default_rng(6).normal(size=(200, 24))
    wi, wd = oracle.knn_double(pts, qs)
    a = tm.AnnKdTree(pts)
    with ThreadPoolExecutor(max_workers=16) as pool:
        got = list(pool.map(lambda i: a.search(qs[i]), range(len(qs))))
    a.destroy()
    assert [g[0] for g in got] == list(wi) and np.allclose([g[1] for g in got], wd, rtol=1e-12)
