// TEST INFRASTRUCTURE — see standin_yarp.h
#pragma once
#include <standin_yarp.h>
