#!/bin/bash
# A/B of the per-level tile pitch (JDA_B200_PITCH_EXTRA) on resident 512-frame batches; by-distribution numbers included.
source tools/ab.sh
EXTRA=" "   # keep the by-distribution breakdown
run default
run p24_16 JDA_B200_PITCH_EXTRA=24:16
run p24_16_37_16 JDA_B200_PITCH_EXTRA=24:16,37:16
run p24_16_30_32_37_16 JDA_B200_PITCH_EXTRA=24:16,30:32,37:16
run p24_80 JDA_B200_PITCH_EXTRA=24:80
