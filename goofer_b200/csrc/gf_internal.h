// gf_internal.h -- declarations shared between the translation units of libgoofer_b200.so
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "../../include/goofer_b200.h"
#include "gf_plan.h"

void gf_set_error(const char *fmt, ...);
int gf_plan_note(const GooferBatch *b, int idx, GfNotePlan *pl);
void gf_plan_info(const GfNotePlan *pl, GooferNotePlanInfo *info);
