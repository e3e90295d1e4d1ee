#!/bin/bash
# One-GPU evidence run of a round: default bench line, reference arm, config-4 / config-5 legs, environment sweep and the ncu
# launch list of the bench command.  tools/collect_round.sh <tag>  -> gpurun_out/<tag>_*.json|csv|jsonl
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err
python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
python bench.py --workload dcp --steps 10 > $out/${tag}_bench_dcp.json 2> $out/${tag}_bench_dcp.err
python bench.py --workload quadrotor --steps 20 > $out/${tag}_bench_quadrotor.json 2> $out/${tag}_bench_quadrotor.err
python bench.py --workload cartpole --steps 10 > $out/${tag}_bench_cartpole.json 2> $out/${tag}_bench_cartpole.err
: > $out/${tag}_env_sweep.jsonl
for c in "PendulumKnown 512 200 20" "PendulumKnown 2048 200 20" "PendulumKnown 4096 200 20" "PendulumKnown 4736 200 20" "PendulumKnown 8192 200 20" \
         "PendulumKnown 16384 200 10" "PendulumKnown 32768 200 10" "PendulumKnown 56832 200 10" "PendulumKnown 65536 200 10" "PendulumKnown 100000 100 10" "PendulumKnown 262144 100 10" \
         "CartpoleKnown 4096 200 10" "CartpoleKnown 8192 200 10" "CartpoleKnown 16384 200 5" "CartpoleKnown 32768 100 5" "DoubleCartpoleKnown 2048 500 4" "DoubleCartpoleKnown 16384 100 4" \
         "DoubleCartpoleKnown 37888 60 4" "DoubleCartpoleKnown 151552 60 3" "Quadrotor 8192 10 20" "Quadrotor 32768 100 4" "Quadrotor 75776 60 3"; do
  set -- $c
  python tools/bench_env.py --env $1 --problems $2 --horizon $3 --iters $4 2>/dev/null | tail -1 >> $out/${tag}_env_sweep.jsonl
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_steps3.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
# one full ncu capture of the throughput kernel (em_ticket_kernel), second matching launch = the timed one
ncu --set full --clock-control none --import-source on -k regex:em_ticket -s 1 -c 1 -f -o $out/${tag}_prof_ticket \
  python tools/bench_env.py --env PendulumKnown --problems 65536 --horizon 100 --iters 5 > $out/${tag}_prof_ticket.log 2>&1
tail -2 $out/${tag}_env_sweep.jsonl
