#!/bin/bash
# Counts the SASS mnemonics that prove tcgen05 / TMEM / TMA / packed-fp32 use, per kernel of the shipped libldx.so
# (B200_PROFILING.md: UTCHMMA = tcgen05.mma fp16, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store,
# UTCBAR = tcgen05.commit, FFMA2 / FMUL2 = packed fp32 pairs; HMMA would be the legacy mma.sync path).
cd "$(dirname "$0")/.."
cuobjdump -sass lightdiffusion_b200/libldx.so | awk '
/Function :/ { name = $3; sub(/^_ZN3ldx[0-9]*/, "", name); }
/UTCHMMA/ { c[name, "UTCHMMA"]++ } /LDTM/ { c[name, "LDTM"]++ } /STTM/ { c[name, "STTM"]++ }
/UTMALDG/ { c[name, "UTMALDG"]++ } /UTMASTG/ { c[name, "UTMASTG"]++ } /UTCBAR/ { c[name, "UTCBAR"]++ }
/FFMA2/ { c[name, "FFMA2"]++ } /FMUL2/ { c[name, "FMUL2"]++ } / HMMA/ { c[name, "HMMA"]++ } /MUFU.EX2/ { c[name, "MUFU.EX2"]++ }
/Function :/ { names[name] = 1 }
END {
  printf "%-72s %8s %6s %6s %8s %8s %7s %6s %6s %9s %5s\n", "kernel", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "FFMA2", "FMUL2", "MUFU.EX2", "HMMA";
  for (n in names) if (c[n, "UTCHMMA"] + c[n, "UTMALDG"] + c[n, "FFMA2"] > 0)
    printf "%-72s %8d %6d %6d %8d %8d %7d %6d %6d %9d %5d\n", substr(n, 1, 72), c[n, "UTCHMMA"], c[n, "LDTM"], c[n, "STTM"], c[n, "UTMALDG"], c[n, "UTMASTG"], c[n, "UTCBAR"], c[n, "FFMA2"], c[n, "FMUL2"], c[n, "MUFU.EX2"], c[n, "HMMA"];
}' | (read -r hdr; echo "$hdr"; sort)
