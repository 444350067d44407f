#!/bin/bash
# GPU run X: timing lines of the heat operator in config 4
set -u
FB_VERBOSE=1 timeout 300 python tools/run_configs.py boussinesq --steps 20 2>&1 | grep -E "heat|config" | tail -9 | cut -c1-600
