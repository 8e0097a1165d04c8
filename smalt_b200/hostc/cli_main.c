/* cli_main.c - entry point of the smalt_b200 executable (everything else lives in
 * libsmalt_b200_map.so so that the same driver is callable in-process, include/smalt_b200_map.h) */
int smalt_b200_cli_main(int argc, char *argv[]);
int main(int argc, char *argv[]) { return smalt_b200_cli_main(argc, argv); }
