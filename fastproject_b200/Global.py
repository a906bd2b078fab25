# -*- coding: utf-8 -*-
"""Seed semantics of ``FastProject/Global.py`` that the hot path depends on.

The reference seeds Python's global ``random`` stream once at import
(/root/reference/FastProject/Global.py:71-72); ``generate_random_sigs`` then
draws from that stream.  The mirror does the same so that, for the same call
history, the random background signatures are the same.
"""
import random

# Chosen by the reference "by roll of a 2147483648-sided die" (Global.py:69-71)
RANDOM_SEED = 1335607828
random.seed(RANDOM_SEED)
