// TEST INFRASTRUCTURE — see standin_idyntree.h
#pragma once
#include <standin_idyntree.h>
