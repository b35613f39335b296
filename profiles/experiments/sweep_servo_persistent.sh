#!/bin/bash
# Sweep of the persistent servo step: tile buffers (1 / 2) x waves of resident CTAs (1 / 2 / 4), with and without
# statistics.  Build here (no GPU needed), run the printed command under gpurun.
#
#   bash profiles/experiments/sweep_servo_persistent.sh build
#   gpurun --timeout 300 -- 'bash profiles/experiments/sweep_servo_persistent.sh run > gpurun_out/servo_sweep.log 2>&1'
#
# `run` first checks every variant's state and statistics against the in-tree build bit for bit (ab_servo_stats.py),
# then times servo_ref / servo_ref+stats / servo_fast+stats at 1,048,576 envs, interleaved in one process.
# A variant that wins still has to pass the GPU suite before it becomes the default:
#   B200CTL_LIB=profiles/experiments/variants/libb200ctl_<name>.so python -m pytest tests/test_gpu_servo.py -m gpu -q
set -e
cd "$(dirname "$0")/../.."
V=profiles/experiments/variants
NAMES=""
for nb in 1 2; do for w in 1 2 4; do for pa in 0 1; do
  [ "$nb$w$pa" = "140" ] && continue          # the in-tree build
  NAMES="$NAMES nb${nb}_w${w}_pa${pa}"
done; done; done
case "$1" in
  build)
    for name in $NAMES; do
      nb=${name:2:1}; w=${name:5:1}; pa=${name:9:1}
      bash profiles/experiments/build_variant.sh "$name" \
        "-DB200_SERVO_STATS_NBUF=$nb -DB200_SERVO_STATS_WAVES=$w -DB200_SERVO_PERSIST_ALL=$pa" | tail -1
    done ;;
  run)
    LIBS=""
    for name in $NAMES; do LIBS="$LIBS $V/libb200ctl_$name.so"; done
    timeout -s KILL 240 python profiles/experiments/ab_servo_stats.py $LIBS ;;
  *) echo "usage: $0 build|run"; exit 2 ;;
esac
