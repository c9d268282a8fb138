#!/bin/bash
# Instruction evidence of the Blackwell-native path: counts of the SASS mnemonics B200_PROFILING.md names, per kernel family,
# from the built library (no GPU needed).   usage: tools/sass_histogram.sh [lib] > profiles/sass_histogram.txt
LIB=${1:-f5_tts_b200/libf5tts_b200.so}
TMP=$(mktemp)
cuobjdump -sass "$LIB" > "$TMP"
echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%d), $(nvcc --version | tail -1))"
echo "# whole library"
for m in UTCHMMA "UTCHMMA.2CTA" LDTM STTM UTMALDG UTMASTG UTMAREDG UTCBAR UBLKCP SYNCS "MUFU.EX2" FFMA2 FADD2 FMNMX3 HGMMA; do
  printf "%-14s %6d\n" "$m" "$(grep -c -- "$m" "$TMP")"
done
printf "%-14s %6d   (legacy mma.sync path: must be 0)\n" "HMMA" "$(grep -c -E '[^C]HMMA' "$TMP")"
echo "# per kernel family (Function headers -> mnemonic counts)"
awk '
  /Function :/ { fn=$0; sub(/.*Function : /,"",fn);
    if (fn ~ /attn_fwd/) fam="attn_fwd_tcgen05_kernel"; else if (fn ~ /gemm_tcgen05/) fam="gemm_tcgen05_kernel (all instantiations)";
    else if (fn ~ /row_norm/) fam="row_norm_kernel"; else if (fn ~ /mel_stft/) fam="mel_stft_kernel"; else if (fn ~ /istft/) fam="istft kernels";
    else fam="other kernels"; nk[fam]++ }
  /UTCHMMA/ {c[fam,"UTCHMMA"]++} /LDTM/ {c[fam,"LDTM"]++} /STTM/ {c[fam,"STTM"]++} /UTMALDG/ {c[fam,"UTMALDG"]++}
  /UTMASTG/ {c[fam,"UTMASTG"]++} /UTMAREDG/ {c[fam,"UTMAREDG"]++} /MUFU.EX2/ {c[fam,"MUFU.EX2"]++} /FFMA2/ {c[fam,"FFMA2"]++}
  /FMNMX3/ {c[fam,"FMNMX3"]++} /HMMA/ && !/UTCHMMA/ {c[fam,"HMMA(legacy)"]++}
  END { n=split("UTCHMMA LDTM STTM UTMALDG UTMASTG UTMAREDG MUFU.EX2 FFMA2 FMNMX3 HMMA(legacy)", ms, " ");
    for (f in nk) { printf "%-44s kernels=%3d", f, nk[f]; for (i=1;i<=n;i++) printf "  %s=%d", ms[i], c[f,ms[i]]+0; printf "\n" } }' "$TMP"
rm -f "$TMP"
