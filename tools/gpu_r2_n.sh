#!/bin/bash
# round-2 pass N (1 GPU): compute-sanitizer passes, then the new push kernel is exercised by a 1-rank communicator
SAN_TIMEOUT=500 bash tools/gpu_sanitize.sh
