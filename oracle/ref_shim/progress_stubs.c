/* The three progress callbacks quantizer.c expects from its host (the reference defines them as no-ops in
 * dlquant/dllmain.c:11-25, a file that cannot be compiled here because it includes Windows.h). */
void progress_init(char *text, int canc) { (void)text; (void)canc; }
int progress_update(float val) { (void)val; return 0; }
void progress_end(void) {}
