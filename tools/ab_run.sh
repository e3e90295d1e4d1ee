#!/bin/bash
# A/B timing of kernel variants on one GPU: tools/ab_run.sh <out.jsonl> <lib1> <lib2> ...   (lib = path or "default";
# append ":ENV=VAL" to set an environment variable for that arm, e.g. default:I2C_B200_NO_HOT=1)
out=$1; shift
: > "$out"
for spec in "$@"; do
  lib=${spec%%:*}; envs=""
  [ "$spec" != "$lib" ] && envs=${spec#*:}
  for cfg in "PendulumKnown 4096 200 20" "PendulumKnown 2048 200 20" "CartpoleKnown 4096 200 10" "PendulumKnown 65536 200 5" "Quadrotor 4096 100 5"; do
    set -- $cfg
    ( [ "$lib" != "default" ] && export I2C_B200_LIB=$lib
      [ -n "$envs" ] && export $envs
      python tools/bench_env.py --env $1 --problems $2 --horizon $3 --iters $4 2>&1 | tail -1 | sed "s|^{|{\"arm\": \"$spec\", |" >> "$out" )
  done
done
cat "$out"
