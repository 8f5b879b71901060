for s in 2 4; do B200_CONV_SLOTS=$s python tools/epi_debug.py; done
