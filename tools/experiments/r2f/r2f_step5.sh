#!/bin/bash
BPP_ACP_TRACE=1 timeout 300 python tools/prof_round.py 52 reference-fixed 4096 16 2>&1 | grep "acp trace" | tail -4
BPP_ACP_TRACE=1 timeout 300 python tools/prof_round.py 52 fixed 4096 16 2>&1 | grep "acp trace" | tail -2
BPP_ACP_TRACE=1 timeout 300 python tools/prof_round.py 4096 fixed 1 8 2>&1 | grep "acp trace" | tail -2
