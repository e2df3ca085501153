#pragma once
#include "sensor.h"
