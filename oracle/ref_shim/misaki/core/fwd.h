// regular.cpp / uniform.cpp include "misaki/core/fwd.h" themselves: same stand-in as for the headers
#pragma once
#include "msk_ref_prelude.h"
