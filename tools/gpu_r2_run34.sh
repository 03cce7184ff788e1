#!/bin/bash
timeout 120 python tools/debug_mref.py 2>&1 | tail -6
