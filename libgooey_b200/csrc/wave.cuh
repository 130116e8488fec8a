// wave.cuh — kernel W: the warp-per-voice back-end.  One warp owns one voice and walks its timeline in blocks of
// (up to) 32 frames, lane = frame.  Within a block
//   * every linear time-invariant recurrence (one-poles, all-pass sections of the half-band oversampler, TPT / Chamberlin
//     state-variable filters, RBJ biquads, DC blocker) is evaluated as a Kogge-Stone prefix scan over the lanes:
//     y[n] = a y[n-1] + b[n] for scalars, s[n] = M s[n-1] + f[n] over 2x2 state matrices for second-order sections;
//   * every memoryless stage between them (tanh waveshaping at 4x rate, gain compensation, sine lookups of phase-
//     modulated oscillators, mixes) is plain SIMT work, one frame per lane;
//   * the few genuinely nonlinear / integer recurrences (xorshift RNGs, f32 phase accumulators, the envelope follower's
//     data-dependent coefficient, the hi-hat's asymmetric smoother) are replayed in the reference's exact operation order
//     over the block (a handful of instructions per frame; every lane runs the same loop and keeps its own frame).
// This turns the serial tail that bounds kernel C (1024 voices = 32 warps on 148 SMs, IPC 0.19) into 1024 warps of mostly
// data-parallel work.  Scans re-associate the f32 sums, so W is not bit-identical to the per-sample order (C / S are);
// the difference is the rounding noise of the recurrence itself (measured <= 3e-6 through the 4x oversampler at drive 40),
// inside the 1e-5 parity bar.  State is the same Aud struct the serial paths use, so calls can alternate between them.
//
// Group width.  The replayed recurrences cost one issue slot per frame for the whole warp, whatever the number of lanes
// that need the result, so a warp may be split into 32/G groups of G lanes, each group owning one voice and a G-frame
// block: replay cost per voice-frame drops by 32/G while the scans get log2(G) steps per G frames instead of 5 per 32.
// The body (wave_impl.inc) is compiled for G = 32, 16 and 8 (namespaces gd::w32 / w16 / w8); pool.cuh picks per voice type.
#pragma once
#include "kernels.cuh"

// Register budget per back end: __launch_bounds__(32, MIN_CTAS).  1 = let ptxas take what it wants.
#ifndef GOOEY_MIN_CTAS_KICK
#define GOOEY_MIN_CTAS_KICK 1
#endif
#ifndef GOOEY_MIN_CTAS_SNARE
#define GOOEY_MIN_CTAS_SNARE 1
#endif
#ifndef GOOEY_MIN_CTAS_HAT
#define GOOEY_MIN_CTAS_HAT 1
#endif
#ifndef GOOEY_MIN_CTAS_TOM
#define GOOEY_MIN_CTAS_TOM 1
#endif

#define WG 32
#define WNS w32
#include "wave_impl.inc"
#undef WG
#undef WNS
// Measured on B200 (C2, r1): G = 32 -> 82 ms / step, 16 -> 105 ms, 8 -> 150 ms.  Narrow groups cut the instruction count
// (snare: 609 M -> 500 M per chunk) but also the warp count (1024 -> 256 per type), and these kernels are bound by the
// dependent-issue latency of each warp's timeline, not by issue slots — so the narrow variants are only compiled on request.
#ifdef GOOEY_WAVE_ALL_WIDTHS
#define WG 16
#define WNS w16
#include "wave_impl.inc"
#undef WG
#undef WNS
#define WG 8
#define WNS w8
#include "wave_impl.inc"
#undef WG
#undef WNS
#endif  // GOOEY_WAVE_ALL_WIDTHS
