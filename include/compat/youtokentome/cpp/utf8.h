/* Drop-in for the reference's youtokentome/cpp/utf8.h (decode_utf8 / encode_utf8, utf8.h:66-80). */
#pragma once
#include "../../../bpe_b200.h"
