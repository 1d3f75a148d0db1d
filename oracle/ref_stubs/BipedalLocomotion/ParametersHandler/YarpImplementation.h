// TEST INFRASTRUCTURE — see standin_blf.h
#pragma once
#include <standin_blf.h>
