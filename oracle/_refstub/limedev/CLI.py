def function_cli(args, name):
    return 0
