#pragma once
#include "common.cuh"
#include "rows.cuh"
