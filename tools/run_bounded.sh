#!/bin/bash
# usage: tools/run_bounded.sh SECONDS LOGFILE cmd...   -- runs cmd in its own session, kills the whole process group
# after SECONDS (a multi-rank launcher's workers die with it, nothing keeps a pipe open).
limit=$1; log=$2; shift 2
setsid "$@" > "$log" 2>&1 &
pid=$!
for ((i = 0; i < limit; i++)); do
  kill -0 $pid 2>/dev/null || break
  sleep 1
done
if kill -0 $pid 2>/dev/null; then
  echo "[run_bounded] killing process group $pid after ${limit}s" >> "$log"
  kill -KILL -- -$pid 2>/dev/null
  sleep 1
  exit 124
fi
wait $pid
