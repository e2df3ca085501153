#pragma once
// Log / Throw come from the prelude (fmt / spdlog are not available)
#include "object.h"
