#pragma once
#include "ubd_handle.cuh"
