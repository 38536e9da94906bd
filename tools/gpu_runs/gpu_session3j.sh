#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 120 python tools/sanitizer_cases.py > $O/s3j_plain.log 2>&1; echo "plain exit $?" >> $O/s3j_plain.log
( timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python tools/sanitizer_cases.py > $O/s3j_racecheck.log 2>&1; echo "racecheck exit $?" >> $O/s3j_racecheck.log )
( timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitizer_cases.py > $O/s3j_synccheck.log 2>&1; echo "synccheck exit $?" >> $O/s3j_synccheck.log )
tail -3 $O/s3j_plain.log
grep -c "Race reported\|hazard" $O/s3j_racecheck.log
tail -4 $O/s3j_racecheck.log
tail -3 $O/s3j_synccheck.log
