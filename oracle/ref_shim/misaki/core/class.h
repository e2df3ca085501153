#pragma once
#include "object.h"
