// stand-in: path.cpp includes <tbb/parallel_for.h> and never uses it
#pragma once
