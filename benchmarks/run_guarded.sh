#!/bin/bash
# run a command with a resident-set watchdog: kills it when the process tree exceeds LIMIT_GB of host memory,
# so that a too-large synthetic refinement cannot take the box down.  usage: run_guarded.sh LIMIT_GB cmd...
limit_kb=$(( $1 * 1000000 )); shift
"$@" &
pid=$!
(
  while kill -0 $pid 2>/dev/null; do
    rss=$(ps -o rss= --ppid $pid -p $pid 2>/dev/null | awk '{s+=$1} END {print s+0}')
    if [ "$rss" -gt "$limit_kb" ]; then echo "run_guarded: rss ${rss} kB over limit, killing $pid" >&2; kill -9 $pid; fi
    sleep 2
  done
) &
wait $pid
