// stand-in: srgb.cpp guards its lazily loaded model with a tbb::spin_mutex; the pinned build is single-threaded
#pragma once
#include "spin_mutex.h"
