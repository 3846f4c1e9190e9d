#!/usr/bin/env bash
# phase table (share of warp samples / instructions) of a fingerprint-kernel ncu report: fp_phase_table.sh REPORT [top]
set -eu
REP=$1; TOP=${2:-30}
F=warpdemux_b200/csrc/fingerprint_kernel.cuh
ncu -i $REP --page source --csv --print-source cuda,sass > /tmp/fp_phase.csv 2>/dev/null
ln() { grep -n "$1" $F | head -1 | cut -d: -f1; }
A=$(ln "^__device__ float block_median_f32_linear"); B=$(ln "^// x / w for a small positive integer")
C=$(ln "^// numpy's pairwise summation"); D=$(ln "^// ---- scipy find_peaks(scores")
E=$(ln "// ---- scipy _select_by_peak_distance"); G=$(ln "// ---- the k_events highest-scoring kept peaks ---"); H=$(ln "^constexpr int FP_RANK_MAX")
I=$(ln "^// ---- consensus sub-sequence match"); K=$(ln "// ---- extract_adapter (sig_proc"); L=$(ln "// NaN padding inside the slice")
M=$(ln "// ---- winsorise at med"); N=$(ln "// ---- c_windowed_t_test (_c_seg"); O=$(ln "// ---- change points: find_peaks")
P=$(ln "// ---- c_new_means (_c_seg"); Q=$(ln "// ---- mean_normalize (sig_proc"); R=$(ln "// ---- statistics (sig_proc"); S=$(ln "// ---- keep the last barcode_num_events (sig")
python scripts/ncu_lines.py /tmp/fp_phase.csv $TOP "median:$A-$((B-1));tthelp:$B-$((C-1));pairwise:$C-$((D-1));localmax:$D-$((E-1));suppress:$E-$((G-1));select:$G-$((H-1));topk:$H-$((I-1));load:$K-$((L-1));params:$L-$((M-1));winsor:$M-$((N-1));ttest:$N-$((O-1));peakcalls:$O-$((P-1));means:$P-$((Q-1));normalize:$Q-$((R-1));stats:$R-$((S-1));out:$S-2000"
