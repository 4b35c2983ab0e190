#!/bin/bash
timeout 600 python scripts/debug_pl.py 2>&1 | grep -v Warning | tail -20 | cut -c1-500
