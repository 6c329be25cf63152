// stream.FindReader kernels (streaming.go:85-255): see capi_stream.inc.
#pragma once
#include "engines.cuh"
