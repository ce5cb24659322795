// Opaque handle layouts behind include/dsm_b200.h (shared by capi.cu and sched.cu).
#pragma once

#include "lm.cuh"
#include "mimi.cuh"

struct dsm_mimi { dsm::MimiImpl* impl; };
struct dsm_lm { dsm::LmImpl* impl; };
struct dsm_asr { dsm::AsrImpl* impl; dsm_mimi* mimi; dsm_lm* lm; };
