#!/bin/bash
# check (self-tests, parity tests, bench) -> evidence profile; stops at the first failure.  Usage: tools/gpu_round.sh <tag>
tag=${1:-r}
tools/gpu_check.sh $tag || exit 1
timeout 400 tools/gpu_profile.sh ${tag}p > /dev/null 2>&1
