#!/bin/bash
scripts/r2_keep4.sh
scripts/r2_keep2.sh
