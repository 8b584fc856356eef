#!/usr/bin/env bash
cd "$(dirname "$0")/.."
timeout 600 python tools/lsq_probe.py 2>&1 | tail -12 | cut -c1-700
