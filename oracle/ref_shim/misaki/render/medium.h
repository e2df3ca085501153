// Stand-in for include/misaki/render/medium.h.  TEST INFRASTRUCTURE.
#pragma once
#include <misaki/core/object.h>
namespace misaki { class Medium : public Object {}; }
