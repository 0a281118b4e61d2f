#!/bin/bash
# Round-2 evidence run on one B200 (under gpurun): GPU tests, the bench line, the ncu launch list of one cfg2 align and
# --set full captures of the hot kernels.  usage: tools/r2_capture.sh <tag> [smoke] [tests] [bench] [launches] [full]
tag=$1; shift
what=" $* "
out=gpurun_out
if [[ $what == *" smoke "* ]]; then
  python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
fi
if [[ $what == *" tests "* ]]; then
  python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1
  tail -3 $out/${tag}_tests.log
fi
if [[ $what == *" bench "* ]]; then
  python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
  python bench.py --impl reference --steps 1 --warmup 0 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
fi
if [[ $what == *" launches "* ]]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $out/${tag}_launches.csv \
      python tools/ncu_align.py 95 > $out/${tag}_ncu_launches.log 2>&1
fi
if [[ $what == *" full "* ]]; then
  for it in 1 45 75 88; do
    ncu --set full --clock-control none --import-source on -k regex:search_kernel -s $it -c 1 -f -o $out/${tag}_search_it$it \
        python tools/ncu_align.py 95 > $out/${tag}_ncu_it$it.log 2>&1
  done
  for it in 45 88; do
    ncu --set full --clock-control none --cache-control none -k regex:search_kernel -s $it -c 1 -f -o $out/${tag}_search_it${it}_warm \
        python tools/ncu_align.py 95 > $out/${tag}_ncu_it${it}_warm.log 2>&1
  done
  ncu --set full --clock-control none --import-source on -k 'regex:trim_moments_kernel|moments_kernel' -s 170 -c 2 -f -o $out/${tag}_trim_moments_it85 \
      python tools/ncu_align.py 95 > $out/${tag}_ncu_tm.log 2>&1
fi
ls -la $out | grep ${tag}_ | head -40
