"""Op/cell registry, decoders and encoder of the reference's src/nn, executed by libnasb200 kernels."""
