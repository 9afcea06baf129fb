/* Drop-in for the reference's youtokentome/cpp/utils.h (Status, BpeConfig, SpecialTokens, BPEState, ... utils.h:11-105). */
#pragma once
#include "../../../bpe_b200.h"
