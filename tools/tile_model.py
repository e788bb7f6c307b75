#!/usr/bin/env python3
"""Static cost model of the 3-D tile kernels (complex single, L taps): per plane-tile FFMA2 warp instructions,
shared-memory / L1 data-pipe wavefronts (128 bytes each) and the three per-launch floors they imply on a B200
(FMA pipe: one FFMA2 warp instruction per 2 cycles per scheduler; L1 data pipe: one wavefront per cycle per SM;
HBM: compulsory bytes at the measured copy peak).  Checked against ncu on cfg5 (profiles/r01_cfg5_final_full_ncu.txt):
k_dec3_fused 559 M shared wavefronts measured vs 554 M modelled; k_rec3_rows 626 M vs 607 M.
The model is what round 1 used to pick experiments; it says nothing about latency hiding (the measured
kernels sit at 45-50 % issue utilisation), which is where the remaining time goes.

usage: tile_model.py [n1 n2 n3 batches]     (default: cfg5's tile launches, 192 192 64 96)"""
import sys

SM, GHZ, PEAK_GBS = 148, 1.955, 6543.4
E = 8            # bytes per element (complex single)
VEC = 2          # elements per 16-byte chunk


def wf(nbytes):
    return nbytes / 128.0


def dec3(L=8, T1=32, T2=16, R2=2, CW=2, R1=4):
    """k_dec3_fused: register ring along dim 3 (stage A), dim 1 in shared memory (B), dim 2 + stores (C)."""
    H = L - 1
    W1, W2 = T1 + H, T2 + H
    npos = W1 * W2
    ffma2 = npos * 2 * L                       # stage A: lo3 / hi3 at every haloed position
    ffma2 += 2 * W2 * T1 * 2 * L               # stage B: 2 arrays x haloed rows x T1 columns x (lo1, hi1)
    ffma2 += 4 * T2 * T1 * 2 * L               # stage C: 4 arrays x (lo2, hi2)
    sts_a = wf(npos * 2 * E)
    nch = (R1 + L - 1 + VEC - 1) // VEC
    items_b = 2 * (T1 // R1) * W2
    lds_b = wf(items_b * nch * 16)
    sts_b = wf(items_b * 2 * R1 * E)
    items_c = 4 * (T2 // R2) * (T1 // CW)
    lds_c = wf(items_c * (R2 + L - 1) * CW * E)
    glob = wf(npos * E) + wf(8 * T1 * T2 * E)   # haloed loads + 8 subband stores
    return dict(voxels=T1 * T2, ffma2_warp=ffma2 / 32, shared_wf=sts_a + lds_b + sts_b + lds_c, global_wf=glob,
                dram_bytes=(1 + 8) * T1 * T2 * E)


def rec3_bulk(L=8, T1=32, T2=16, R2=8, R1=4):
    """k_rec3_bulk: TMA-staged haloed tiles (W1S x W2 per band), dim 2 (RA), dim 1 (RB), scatter ring (RC)."""
    H = L - 1
    W1, W2, W1S = T1 + H, T2 + H, T1 + 8
    ffma2 = 4 * T2 * W1 * 2 * L + 2 * T2 * T1 * 2 * L + T1 * T2 * 2 * L
    tma = wf(8 * W2 * W1S * E)
    lds_a = wf(4 * (T2 // R2) * W1 * 2 * (R2 + L - 1) * E)
    sts_a = wf(4 * T2 * W1 * E)
    nch = (R1 + L - 1 + VEC - 1) // VEC
    items_b = 2 * (T1 // R1) * T2
    lds_b = wf(items_b * 2 * nch * 16)
    sts_b = wf(2 * T2 * T1 * E)
    lds_c = wf(2 * T2 * T1 * E)
    return dict(voxels=T1 * T2, ffma2_warp=ffma2 / 32, shared_wf=lds_a + sts_a + lds_b + sts_b + lds_c, tma_wf=tma,
                global_wf=wf(T1 * T2 * E), dram_bytes=(8 + 1) * T1 * T2 * E,
                l2_read_bytes=8 * W2 * W1S * E)


def rec3_rows(L=8, n1=192, T2=8, RH=4, R1=4):
    """k_rec3_rows: full rows, band-pair stages, RA on row halves, RB with R1 outputs per item."""
    W2 = T2 + L - 1
    ffma2 = 4 * T2 * n1 * 2 * L + 2 * T2 * n1 * 2 * L + n1 * T2 * 2 * L
    tma = wf(8 * W2 * n1 * E)
    lds_a = wf(4 * (T2 // RH) * n1 * 2 * (RH + L - 1) * E)
    sts_a = wf(4 * T2 * (n1 + L - 1) * E)
    nch = (R1 + L - 1 + VEC - 1) // VEC
    items_b = 2 * (n1 // R1) * T2
    lds_b = wf(items_b * 2 * nch * 16)
    sts_b = wf(2 * T2 * n1 * E)
    lds_c = wf(2 * T2 * n1 * E)
    return dict(voxels=n1 * T2, ffma2_warp=ffma2 / 32, shared_wf=lds_a + sts_a + lds_b + sts_b + lds_c, tma_wf=tma,
                global_wf=wf(n1 * T2 * E), dram_bytes=(8 + 1) * n1 * T2 * E, l2_read_bytes=8 * W2 * n1 * E)


def report(name, m, nvox_launch):
    tiles = nvox_launch / m["voxels"]
    per_sm = tiles / SM
    cyc = per_sm * GHZ * 1e6                                    # cycles per ms
    fma_ms = per_sm * m["ffma2_warp"] * 2 / 4 / (GHZ * 1e6)
    pipe_wf = m["shared_wf"] + m["global_wf"] + m.get("tma_wf", 0.0)
    lsu_ms = per_sm * pipe_wf / (GHZ * 1e6)
    hbm_ms = tiles * m["dram_bytes"] / (PEAK_GBS * 1e6)
    print("%-34s voxels/tile %5d  FFMA2 warp-instr %6.0f  shared wf %6.0f (+TMA %4.0f, global %4.0f)  floors: FMA %.2f ms, "
          "L1 data pipe %.2f ms, HBM %.2f ms" % (name, m["voxels"], m["ffma2_warp"], m["shared_wf"], m.get("tma_wf", 0.0),
                                                 m["global_wf"], fma_ms, lsu_ms, hbm_ms))
    return tiles


def main():
    n1, n2, n3, nb = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (192, 192, 64, 96)
    nvox = n1 * n2 * n3 * nb
    print("tile launches over %dx%dx%d x %d batches = %.1f M voxels (complex single, db4)" % (n1, n2, n3, nb, nvox / 1e6))
    t = report("k_dec3_fused 32x16 R2=2 (default)", dec3(), nvox)
    print("   modelled shared wavefronts per launch: %.0f M" % (t * dec3()["shared_wf"] / 1e6))
    report("k_dec3_fused CW=1 R2=8", dec3(R2=8, CW=1), nvox)
    report("k_dec3_fused CW=1 R2=16 R1=8", dec3(R2=16, CW=1, R1=8), nvox)
    report("k_rec3_bulk 32x16 R2=8", rec3_bulk(), nvox)
    report("k_rec3_bulk 32x16 full height R1=8", rec3_bulk(R2=16, R1=8), nvox)
    t = report("k_rec3_rows T2=8 halves R1=4", rec3_rows(n1=n1), nvox)
    print("   modelled shared wavefronts per launch: %.0f M" % (t * rec3_rows(n1=n1)["shared_wf"] / 1e6))
    report("k_rec3_rows T2=8 full height R1=8", rec3_rows(n1=n1, RH=8, R1=8), nvox)


if __name__ == "__main__":
    main()
