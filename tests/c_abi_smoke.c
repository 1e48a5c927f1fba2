#include <stdio.h>
#include "juicy_batch.h"
int main(void) {
    int chain[3] = { JB_PUNCH, JB_TEXTURE, JB_WIDTH };
    jb_engine* e = 0; jb_wav_info info; jb_meter_stats stats[4]; float* audio = 0; size_t sz = 0;
    if (jb_create(chain, 3, 4, 2, -1, &e) != JB_OK) { printf("%s\n", jb_last_error()); return 1; }
    jb_set_param_clips(e, 1, "material", 2.0f, 1, 2);
    jb_set_program_clips(e, 2, 3, 0, 2);
    jb_schedule_param(e, 2, "width", 940, 0.2f, JB_ALL_CLIPS, 0);
    jb_get_state(e, 0, JB_ALL_CLIPS, 0, 0, &sz);
    printf("sets %d state bytes %zu math %d\n", jb_num_param_sets(e), sz, jb_set_math_mode(e, JB_MATH_EXACT));
    (void) info; (void) stats; (void) audio;
    jb_destroy(e);
    return 0;
}
