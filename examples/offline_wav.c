/* Plain-C consumer of libpfasr: WAV files in, text out - what AliParaformerAsr.Examples/Program.cs does through the C#
 * OfflineRecognizer (GetFileSample -> AddSamples -> GetResults), here straight against include/pf_abi.h.
 *
 *   gcc -std=c99 -Iinclude examples/offline_wav.c -Laliparaformerasr_b200 -lpfasr -Wl,-rpath,$PWD/aliparaformerasr_b200 -o offline_wav
 *   ./offline_wav model.pfw tokens.txt a.wav [b.wav ...]
 *   PFASR_EXAMPLE_LAYERS=3,2 ./offline_wav ...      (encoder, decoder layer counts of a smaller model, e.g. the test model)
 *
 * The model description below is paraformer-large (asr.yaml of the published model); a real application fills pf_config
 * from its configuration file the way OfflineModel / ConfEntity do. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pf_abi.h"

static void* read_file(const char* path, size_t* bytes) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void* p = malloc(n > 0 ? (size_t)n : 1);
    if (p && fread(p, 1, (size_t)n, f) != (size_t)n) { free(p); p = NULL; }
    fclose(f);
    *bytes = (size_t)n;
    return p;
}

static int fail(const char* what) {
    fprintf(stderr, "%s: %s\n", what, pf_last_error());
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s model.pfw tokens.txt file.wav [...]\n", argv[0]);
        return 2;
    }
    pf_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_bytes = (int32_t)sizeof cfg;
    cfg.model_kind = PF_MODEL_PARAFORMER;
    cfg.input_size = 560; cfg.d_model = 512; cfg.heads = 4; cfg.ffn = 2048; cfg.enc_layers = 50; cfg.enc_kernel = 11;
    cfg.dec_layers = 16; cfg.dec_ffn = 2048; cfg.dec_kernel = 11; cfg.vocab = 8404;
    cfg.ln_eps = 1e-12f; cfg.cif_threshold = 1.0f; cfg.cif_tail = 0.45f; cfg.smooth_factor = 1.0f; cfg.noise_threshold = 0.0f;
    cfg.fs = 16000; cfg.n_mels = 80; cfg.lfr_m = 7; cfg.lfr_n = 6;
    if (getenv("PFASR_EXAMPLE_LAYERS")) {
        int e = 0, d = 0;
        if (sscanf(getenv("PFASR_EXAMPLE_LAYERS"), "%d,%d", &e, &d) == 2) { cfg.enc_layers = e; cfg.dec_layers = d; }
    }

    pf_tokens* tokens = NULL;
    if (pf_tokens_create(argv[2], &tokens) != PF_OK) return fail("tokens");
    pf_offline* h = NULL;
    if (pf_offline_create(&cfg, argv[1], NULL, 0, &h) != PF_OK) { pf_tokens_destroy(tokens); return fail("create"); }

    const int batch = argc - 3;
    pf_audio* clips = (pf_audio*)calloc((size_t)batch, sizeof(pf_audio));
    void** files = (void**)calloc((size_t)batch, sizeof(void*));
    int rc = 0;
    for (int i = 0; i < batch && rc == 0; ++i) {
        size_t bytes = 0;
        files[i] = read_file(argv[3 + i], &bytes);
        if (!files[i]) { fprintf(stderr, "cannot read %s\n", argv[3 + i]); rc = 1; break; }
        if (pf_wav_parse(files[i], bytes, &clips[i]) != PF_OK) rc = fail(argv[3 + i]);
    }
    pf_result res;
    if (rc == 0 && pf_offline_run_audio(h, clips, batch, 0, &res) != PF_OK) rc = fail("run");
    for (int i = 0; i < batch && rc == 0; ++i) {
        pf_text_result t;
        memset(&t, 0, sizeof t);
        if (pf_decode_offline_result(tokens, &res, i, &t) != PF_OK) { rc = fail("decode"); break; }   /* sizes */
        t.text = (char*)malloc(t.text_bytes + 1);
        t.text_capacity = t.text_bytes + 1;
        if (pf_decode_offline_result(tokens, &res, i, &t) != PF_OK) rc = fail("decode");
        else printf("%s\t%s\n", argv[3 + i], t.text);
        free(t.text);
    }
    for (int i = 0; i < batch; ++i) free(files[i]);
    free(files);
    free(clips);
    pf_offline_destroy(h);
    pf_tokens_destroy(tokens);
    return rc;
}
