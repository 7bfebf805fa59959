#!/bin/bash
timeout 400 python examples/vit_gamma_recipe.py 2>&1 | grep -v Warning | tail -20
